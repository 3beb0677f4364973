// Fused rela-fusion layer for sm_100a: TMA-staged edge tiles, tcgen05.mma (kind::f16, fp32
// accumulators in TMEM) for the four N^2 contractions, warp-level LayerNorm / online softmax
// epilogues straight out of TMEM.  Reference semantics: RelaFusionLayer._build_memory and
// _mha_block, planners/mind/networks/network.py:182-226.
//
// Work decomposition: one work item = (scene b, 16 queries j0..j0+15); the CTA walks the keys in
// chunks of 8, so a tile is 8 keys x 16 queries = 128 pair rows = the 128 TMEM lanes.  A scene's last
// query block, when it holds a single query (n % 16 == 1: the cls token of a 32 x 128 scene), is a
// single-query item instead: tiles of 128 keys x 1 query (tc_build_schedule).
//   G1:  D1[128x128]  = edge_tile(fp16) . W_e^T                      (+S[j]+T[i], LN, ReLU -> memory)
//   G2:  Dpe|Dk|Dv    = memory(fp16)   . [W_pe ; W_k ; W_v]^T
//   edge' = LN(edge + ReLU(LN(Dpe + b)))  written back in place (fp16) by TMA
//   per-thread online softmax over the keys a thread sees, merged across the 8 key slots at the
//   end of the work item.  `memory`, K and V never touch HBM.
#include "fusion_tc.h"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

namespace mind {

namespace tc {

constexpr int kThreads = 544;   // 16 epilogue warps + 1 issuer warp
constexpr int kIssuerWarp = 0;  // warps 1..16 are the epilogue warps
constexpr float kEps = 1e-5f;

// ---- shared memory map (bytes, relative to a 1024-aligned base) ----
constexpr uint32_t SM_W = 0;                          // [2 kblock][512 rows][128 B]  = 131072
constexpr uint32_t SM_TILE0 = 131072;                 // [2 kblock][128 rows][128 B]  = 32768
constexpr uint32_t SM_TILE1 = SM_TILE0 + 32768;
constexpr uint32_t SM_S = SM_TILE1 + 32768;           // float [16][132]
constexpr uint32_t SM_Q = SM_S + 16 * 132 * 4;        // float [16][132]
constexpr uint32_t SM_T = SM_Q + 16 * 132 * 4;        // float [8][128], written by TMA (dense rows)
constexpr uint32_t SM_P = SM_T + 8 * 128 * 4;         // float [8][128] layer params
constexpr uint32_t SM_STAT = SM_P + 8 * 128 * 4;      // float2 [2 buf][4 quarter][128]
constexpr uint32_t SM_BAR = SM_STAT + 2 * 4 * 128 * 8;
constexpr uint32_t SM_TMEM = SM_BAR + 96;   // 10 mbarriers
constexpr uint32_t SM_TOTAL = SM_TMEM + 16;
constexpr uint32_t SMEM_BYTES = SM_TOTAL + 1024;      // slack for manual 1024 B alignment
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");

// param rows
enum { P_MEM_G = 0, P_MEM_B, P_BPE, P_PE_G, P_PE_B, P_NE_G, P_NE_B, P_BV };

// error codes written to *err before trapping
enum { E_LOAD_W = 1, E_LOAD_EDGE = 2, E_MMA1 = 3, E_MMA2 = 4 };

using namespace mind::tcp;      // PTX wrappers: tc_ptx.cuh

// ---------------------------------------------------------------------------------------------
// optional timeline trace (build with -DMIND_TRACE -> libmind_b200_trace.so): CTA 0 records clock64()
// at the hand-off points of tiles [kTraceG0, kTraceG0 + kTraceTiles); read back with mind_trace_read
// ---------------------------------------------------------------------------------------------
#ifdef MIND_TRACE
#ifndef MIND_TRACE_G0
#define MIND_TRACE_G0 40
#endif
constexpr int kTraceG0 = MIND_TRACE_G0, kTraceTiles = 8, kTracePts = 32;
__device__ long long g_trace[kTraceTiles * 17 * kTracePts];
#define TR(id, gg)                                                                                        \
    do {                                                                                                  \
        if (blockIdx.x == 0 && a.has_edge && (threadIdx.x & 31) == 0 && (int)(gg) >= kTraceG0 && (int)(gg) < kTraceG0 + kTraceTiles) \
            g_trace[(((int)(gg) - kTraceG0) * 17 + (((threadIdx.x >> 5) + 16) % 17)) * kTracePts + (id)] = clock64();  \
    } while (0)
#elif defined(MIND_PROGRESS)
// deadlock finder: every warp posts (tile << 8 | point) into a mapped host array that survives a trapped launch
__device__ volatile int* g_progress;
#define TR(id, gg)                                                                                        \
    do {                                                                                                  \
        if ((threadIdx.x & 31) == 0) g_progress[blockIdx.x * 17 + (threadIdx.x >> 5)] = ((int)(gg) << 8) | (id); \
    } while (0)
#else
#define TR(id, gg) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// the fused layer kernel
// ---------------------------------------------------------------------------------------------
struct LayerArgs {
    const TcWork* work;
    int n_work;
    const float* stq;     // [B*Nmax, 384]
    const float* params;  // [8][128]
    __half* attn_hi;      // [B*Nmax, 128] attention output as an fp16 (hi, lo) pair: A operand of the out-proj GEMM
    __half* attn_lo;
    int Nmax;
    int has_edge;
    int* err;
    float* part;          // partial softmax states of key-split items: [slot][16 j][144]
};

// TMEM column map (512 columns x 128 lanes, fp32 cells)
//   [  0,128)  D1 = edge.W_e^T ; after epilogue 1 the same columns hold the A operand of G2:
//              memory as an fp16 (hi, lo) pair, two K elements per 32-bit cell: hi -> [0,64), lo -> [64,128)
//   [128,256)  Dpe   [256,384)  Dk   [384,512)  Dv
// Roles: warps 0-15 = epilogue (warp w owns TMEM lanes 32*(w&3).. = pair rows, and the 32-channel
// column quarter q = w>>2); warp 16 = issuer (TMA loads / stores, every tcgen05.mma, T-tile staging).
// Steady state has no CTA-wide barrier: issuer and epilogue meet only through mbarriers, the four
// warps that share rows through a named 128-thread barrier.
struct TileIt {          // walks (work item, key chunk) over the CTA's contiguous range [wi, wi_end) of the work list
    int wi, wi_end, ch, ch1, b, j0, mode;
};
__device__ __forceinline__ bool tile_valid(const TileIt& t) { return t.wi < t.wi_end; }
__device__ __forceinline__ void tile_load_wi(TileIt& t, const TcWork* work) {
    if (t.wi < t.wi_end) {
        const TcWork w = work[t.wi];
        t.b = w.b; t.j0 = w.j0; t.ch = w.ch0; t.ch1 = w.ch1; t.mode = w.mode;
    }
}
__device__ __forceinline__ void tile_next(TileIt& t, const TcWork* work) {
    if (++t.ch >= t.ch1) { ++t.wi; tile_load_wi(t, work); }
}

__global__ void __launch_bounds__(kThreads, 1)      // 17 warps are granted registers as 5 warpgroups -> 96 / thread
k_rela_fusion_tc(const __grid_constant__ CUtensorMap emap, const __grid_constant__ CUtensorMap emapq,
                 const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap tmap, LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));   // generic pointer to the aligned base
    float* sS = reinterpret_cast<float*>(sgen + SM_S);
    float* sQ = reinterpret_cast<float*>(sgen + SM_Q);
    float* sT = reinterpret_cast<float*>(sgen + SM_T);
    float* sP = reinterpret_cast<float*>(sgen + SM_P);
    float2* sStat = reinterpret_cast<float2*>(sgen + SM_STAT);   // [2 buf][4 quarter][128 row]
    volatile uint32_t* sTmem = reinterpret_cast<volatile uint32_t*>(sgen + SM_TMEM);
    const uint32_t bar_w = sbase + SM_BAR;
    const uint32_t bar_m1 = bar_w + 8, bar_m2a = bar_w + 16, bar_m2b = bar_w + 24, bar_ld0 = bar_w + 32;   // +32, +40
    const uint32_t bar_a = bar_w + 48, bar_e = bar_w + 56, bar_t = bar_w + 64, bar_k = bar_w + 72;

    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it (TMEM
    // addresses, barrier ids, role tests) in uniform registers instead of an R2UR in front of every tcgen05.ld / st
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_m1, 1); mbar_init(bar_m2a, 1); mbar_init(bar_m2b, 1);
        mbar_init(bar_ld0, 1); mbar_init(bar_ld0 + 8, 1);
        mbar_init(bar_a, 16); mbar_init(bar_e, 16); mbar_init(bar_t, 1); mbar_init(bar_k, 16);
        fence_barrier_init();
    }
    if (warp == kIssuerWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_TMEM), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 8 * 128; i += kThreads) sP[i] = a.params[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    // work[0 .. gridDim.x) are range headers: CTA c owns the work items [hdr.b, hdr.j0) (tc_prepare)
    const int wi_begin = a.work[blockIdx.x].b, wi_end = a.work[blockIdx.x].j0;

    if (warp == kIssuerWarp) {
        // =============================== issuer warp ===============================
        if (elect_one()) {   // weights: 8 boxes of [128 rows x 64 k] -> SM_W[kblock][row]
            mbar_expect_tx(bar_w, 131072u);
            for (int kb = 0; kb < 2; ++kb)
                for (int m = 0; m < 4; ++m) tma_load_2d(sbase + SM_W + kb * 65536 + m * 16384, &wmap, bar_w, kb * 64, m * 128);
            mbar_wait(bar_w, 0, a.err, E_LOAD_W);
        }
        __syncwarp();
        auto load_tile = [&](const TileIt& t, int buf) {          // lane 0 only
            const uint32_t bl = bar_ld0 + 8 * buf, dst = sbase + SM_TILE0 + buf * 32768;
            mbar_expect_tx(bl, 32768u);
            if (t.mode) {     // single-query item: box [64 c][1 j][128 i], tile row = key
                tma_load_4d(dst, &emapq, bl, 0, t.j0, t.ch * 128, t.b);
                tma_load_4d(dst + 16384, &emapq, bl, 64, t.j0, t.ch * 128, t.b);
            } else {
                tma_load_4d(dst, &emap, bl, 0, t.j0, t.ch * 8, t.b);
                tma_load_4d(dst + 16384, &emap, bl, 64, t.j0, t.ch * 8, t.b);
            }
        };
        // T rows (tar term, per key i) of a tile: one TMA box [128 floats x 8 token rows] out of stq[:, 128:256).  Rows of
        // padding tokens (i >= n) carry finite values and are masked in the softmax; rows past the tensor are zero-filled.
        // Single-query items read their T rows from global memory; the box is still loaded to keep the protocol uniform.
        auto load_T = [&](const TileIt& t) {                      // lane 0 only
            mbar_expect_tx(bar_t, 4096u);
            tma_load_2d(sbase + SM_T, &tmap, bar_t, 128, t.b * a.Nmax + (t.mode ? 0 : t.ch * 8));
        };
        auto issue_g1 = [&](int buf, uint32_t dcol) {             // lane 0 only
            const uint32_t tX = sbase + SM_TILE0 + buf * 32768, id128 = umma_idesc_f16(128);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t ko = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                const uint32_t kw = (uint32_t)(kk >> 2) * 65536u + (uint32_t)(kk & 3) * 32u;
                umma_f16(tmem + dcol, umma_desc_sw128(tX + ko), umma_desc_sw128(sbase + SM_W + kw), id128, kk > 0);
            }
            umma_commit(bar_m1);
        };
        TileIt cur_t{wi_begin, wi_end, 0, 0, 0, 0, 0};
        tile_load_wi(cur_t, a.work);
        if (tile_valid(cur_t)) {
            TileIt nxt = cur_t;
            tile_next(nxt, a.work);
            uint32_t g = 0;                                       // tiles processed by this CTA
            if (elect_one()) {
                load_tile(cur_t, 0);
                if (tile_valid(nxt)) load_tile(nxt, 1);
            }
            if (elect_one()) {
                load_T(cur_t);
                mbar_wait(bar_ld0, 0, a.err, E_LOAD_EDGE);
                tc_fence_after();
                issue_g1(0, 0u);
            }
            while (true) {
                const int buf = g & 1;
                const bool has_next = tile_valid(nxt);
                // Layer without edge update: the W_pe accumulator columns are free, so D1 / the A operand alternate between
                // columns [0,128) and [128,256) from tile to tile.  G1 of the next tile can then be issued AHEAD of this
                // tile's K|V products (which still read this tile's A operand) and epilogue 1 of the next tile runs under
                // them; with a single D1 region that layer was bound by the serial chain E1 -> K|V MMAs -> G1 -> E1.
                const uint32_t acol = (!a.has_edge && (g & 1)) ? 128u : 0u;
                const uint32_t ncol = (!a.has_edge && !(g & 1)) ? 128u : 0u;
                // [A] A operand of this tile is in TMEM -> G2 (W_pe first: its epilogue overlaps the K|V MMAs)
                TR(15, g);
                handoff_sync(kBarA);
                TR(16, g);
                if (elect_one()) {
                    tc_fence_after();
                    const uint32_t id128 = umma_idesc_f16(128);
                    if (a.has_edge) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t kw = (uint32_t)(kk >> 2) * 65536u + (uint32_t)(kk & 3) * 32u;
                            const uint64_t bd = umma_desc_sw128(sbase + SM_W + kw + 128 * 128);
                            // hi part of `memory` only: for the scenes this kernel serves (>= 128 tokens) the lo term of the
                            // W_pe product is below the noise of the fp16 edge stream (oracle/format_experiments.py)
                            umma_f16_ts(tmem + 128, tmem + kk * 8, bd, id128, kk > 0);
                        }
                    }
                    umma_commit(bar_m2a);
                    if (!a.has_edge && has_next) {                 // [B'] layer without edge update: next tile's T rows and G1 right here
                        // (its D1 region is the other one, free since the previous tile's K|V products were issued; every epilogue
                        // warp is past epilogue 1 of this tile, so the T buffer is free): epilogue 1 of the next tile then never
                        // waits for G1, and it runs under this tile's K|V products
                        load_T(nxt);
                        mbar_wait(bar_ld0 + 8 * (buf ^ 1), ((g + 1) >> 1) & 1, a.err, E_LOAD_EDGE);
                        tc_fence_after();
                        issue_g1(buf ^ 1, ncol);
                    }
                }
                TR(17, g);
                handoff_sync(kBarK);                               // Dk/Dv of the previous tile have been consumed
                TR(18, g);
                if (elect_one()) {
                    tc_fence_after();
                    // K and V as two N=128 groups: with the A operand in TMEM an N=256 MMA measured ~170 cycles against
                    // ~60 for N=128 (timeline trace), so 32 narrow MMAs finish well before 16 wide ones
                    const uint32_t id128 = umma_idesc_f16(128);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t kw = (uint32_t)(kk >> 2) * 65536u + (uint32_t)(kk & 3) * 32u;
                            const uint64_t bd = umma_desc_sw128(sbase + SM_W + kw + (256 + 128 * m) * 128);
                            umma_f16_ts(tmem + 256 + 128 * m, tmem + acol + kk * 8, bd, id128, kk > 0);
                            umma_f16_ts(tmem + 256 + 128 * m, tmem + acol + 64 + kk * 8, bd, id128, 1);
                        }
                    }
                    umma_commit(bar_m2b);
                    TR(19, g);
                    // [B] next tile: T rows (every epilogue warp is past epilogue 1 of this tile), then G1 behind G2
                    if (has_next && a.has_edge) {
                        load_T(nxt);
                        mbar_wait(bar_ld0 + 8 * (buf ^ 1), ((g + 1) >> 1) & 1, a.err, E_LOAD_EDGE);
                        TR(20, g);
                        tc_fence_after();
                        issue_g1(buf ^ 1, ncol);
                        TR(21, g);
                    }
                }
                // [C] edge' tile complete -> TMA store ; [D] buffer free -> load the tile after next
                TileIt nn = nxt;
                if (has_next) tile_next(nn, a.work);
                handoff_sync(kBarE);
                TR(22, g);
                if (elect_one()) {
                    if (a.has_edge) {
                        const uint32_t tX = sbase + SM_TILE0 + buf * 32768;
                        if (cur_t.mode) {
                            tma_store_4d(&emapq, tX, 0, cur_t.j0, cur_t.ch * 128, cur_t.b);
                            tma_store_4d(&emapq, tX + 16384, 64, cur_t.j0, cur_t.ch * 128, cur_t.b);
                        } else {
                            tma_store_4d(&emap, tX, 0, cur_t.j0, cur_t.ch * 8, cur_t.b);
                            tma_store_4d(&emap, tX + 16384, 64, cur_t.j0, cur_t.ch * 8, cur_t.b);
                        }
                        tma_commit();
                    }
                    if (has_next && tile_valid(nn)) {
                        tma_wait_read0();
                        load_tile(nn, buf);
                    }
                }
                if (!has_next) break;
                cur_t = nxt; nxt = nn; ++g;
            }
            if (elect_one()) tma_wait_all0();
        }
    } else {
        // =============================== epilogue warps ===============================
        const int ew = warp - 1;                     // epilogue warp index 0..15
        const int q = ew >> 2;                       // column quarter: channels [32q, 32q+32)
        const int lg = warp & 3;                     // TMEM lane quadrant is fixed by the hardware warp id
        const int row = lg * 32 + lane;              // TMEM lane = pair row inside the tile
        const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
        const int col0 = q * 32;
        const uint32_t tile_off = (uint32_t)(q >> 1) * 16384u;      // k-block of this quarter inside an edge tile
        const int chunk0 = (q & 1) * 4;                              // first 16-byte chunk of this quarter in its k-block
        const float* Pm = sP;
        uint32_t g = 0;                                              // tiles processed by this CTA
        for (int wi = wi_begin; wi < wi_end; ++wi) {
            const TcWork wk = a.work[wi];
            const int N = wk.n, j0 = wk.j0, b = wk.b;
            const int ch0 = wk.ch0, ch1 = wk.ch1;
            // tile row -> (key slot, query): 8 keys x 16 queries, or (single-query item) 128 keys x query j0
            const int mode = wk.mode;
            const int i_l = mode ? row : (row >> 4), j_l = mode ? 0 : (row & 15);
            const int kstep = mode ? 128 : 8;
            const int64_t tok0 = (int64_t)b * a.Nmax;
            {   // S (src term, per query j) and q tiles: 16 x 32 float4 each, one per epilogue thread
                const int jj = ew, c4 = tid & 31;
                float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), qv = sv;
                if (j0 + jj < N) {
                    const float* p = a.stq + (tok0 + j0 + jj) * 384;
                    sv = reinterpret_cast<const float4*>(p)[c4];
                    qv = reinterpret_cast<const float4*>(p + 256)[c4];
                }
                *reinterpret_cast<float4*>(sS + jj * 132 + c4 * 4) = sv;
                *reinterpret_cast<float4*>(sQ + jj * 132 + c4 * 4) = qv;
            }
            TR(24, g);
            asm volatile("bar.sync 5, 512;" ::: "memory");         // epilogue warps only
            TR(25, g);
            f2 acc2[16];                                           // un-normalised attention accumulators, 2 heads x 16 channels
            float mrun[2], lrun[2];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc2[k] = 0ull;
            mrun[0] = mrun[1] = -INFINITY; lrun[0] = lrun[1] = 0.f;

            // ---- epilogue 2b: per-thread online softmax over this thread's key, 2 heads.  Runs one tile late
            // (right after the A operand of the NEXT tile is handed over) so that it overlaps the W_pe MMAs.
            // The running maximum is only a reference point: it is moved (and the state rescaled) when a score
            // exceeds it by more than 8, so the common case is one FFMA2 per channel pair.
            auto attend = [&](uint32_t tile_par, bool key_ok) {
                mbar_wait(bar_m2b, tile_par, a.err, E_MMA2);
                TR(5, g);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t rv[16];
                    float s;
                    {
                        uint32_t rk[16];
                        TMEM_LD_X16_NM(tmem + lane_base + 256 + col0 + h * 16, rk);   // warp-collective: no lane guard
                        TMEM_LD_X16_NM(tmem + lane_base + 384 + col0 + h * 16, rv);
                        TMEM_WAIT_LD_R16(rk);
                        TMEM_PIN_R16(rv);
                        f2 sa = 0ull, sb = 0ull;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            f2 qa, qb;
                            lds_2f2(sQ + (j_l * 132 + col0 + h * 16 + k4 * 4), qa, qb);
                            sa = fma2(qa, pk2u(rk[k4 * 4 + 0], rk[k4 * 4 + 1]), sa);
                            sb = fma2(qb, pk2u(rk[k4 * 4 + 2], rk[k4 * 4 + 3]), sb);
                        }
                        s = hsum2(add2(sa, sb));
                    }
                    const bool need = key_ok && (s > mrun[h] + 8.f);              // first valid key: mrun = -inf
                    if (__any_sync(0xffffffffu, need)) {
                        const float corr = need ? __expf(mrun[h] - s) : 1.f;      // exp(-inf) = 0 on the first key
                        if (need) mrun[h] = s;
                        lrun[h] *= corr;
                        const f2 c2 = pk2(corr, corr);
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc2[h * 8 + k] = mul2(acc2[h * 8 + k], c2);
                    }
                    if (key_ok) {
                        const float p = __expf(s - mrun[h]);
                        lrun[h] += p;
                        const f2 p2 = pk2(p, p);
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc2[h * 8 + k] = fma2(p2, pk2u(rv[2 * k], rv[2 * k + 1]), acc2[h * 8 + k]);
                    }
                }
                tc_fence_before();
            };

            for (int ch = ch0; ch < ch1; ++ch, ++g) {
                const int i0 = ch * kstep;
                const uint32_t par = g & 1;
                // single-query item: this row's T (tar) vector comes straight from stq (row clamped: keys past the scene
                // are masked in the softmax, their values only have to be finite)
                const float* Tg = a.stq + (tok0 + min(i0 + i_l, a.Nmax - 1)) * 384 + 128;
                const uint32_t tX = sbase + SM_TILE0 + (g & 1) * 32768;  // edge tile (in place -> edge')
                TR(0, g);
                mbar_wait(bar_t, par, a.err, E_LOAD_EDGE + 10);          // T rows staged by the issuer
                mbar_wait(bar_m1, par, a.err, E_MMA1);
                TR(1, g);
                tc_fence_after();

                // ---- epilogue 1: memory = ReLU(LN(D1 + S[j] + T[i])) -> fp16 (hi, lo) A operand in TMEM ----
                // streaming 16-column passes; x = D1 + S + T is parked in this thread's own Dpe cells
                // (free until G2 of this tile is issued) so nothing has to live in registers across the barrier
                const uint32_t scr_t = tmem + lane_base + 128 + col0;
                const uint32_t d1c = (!a.has_edge && (g & 1)) ? 128u : 0u;      // D1 / A operand region of this tile (see the issuer)
                // the two statistics buffers swap roles from tile to tile (this tile: epilogue 1 and pass B -> sx, pass A -> sy).
                // Every re-use of a buffer is then already ordered by one of the three row-group barriers of the tile body:
                // sy of this tile was last read in pass B of the previous tile (before its third barrier), sx in its pass C
                // (before this tile's first barrier), so no barrier is needed at the end of a tile.
                const int sx = (int)(g & 1), sy = sx ^ 1;
                // A/B build for the next round: x stays in 32 registers across the row-group barrier instead of a round trip
                // through the TMEM scratch cells (2 tcgen05.st + wait::st + 2 tcgen05.ld per thread-tile).  Not yet run on hardware.
                uint32_t xr[32];
                auto e1_pass1 = [&](auto qtag) {                   // two straight-line copies: no branch inside the unrolled loops
                    constexpr bool kQ = decltype(qtag)::value;
                    f2 s1a = 0ull, s1b = 0ull, s2a = 0ull, s2b = 0ull;
                    uint32_t rr[2][16];
                    TMEM_LD_X16_NM(tmem + lane_base + d1c + col0, rr[0]);
                    TMEM_LD_X16_NM(tmem + lane_base + d1c + col0 + 16, rr[1]);
                    TMEM_WAIT_LD_R16(rr[0]);
                    TMEM_PIN_R16(rr[1]);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t (&r)[16] = rr[hf];
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const int c = col0 + hf * 16 + k4 * 4;
                            f2 sa, sb, ta, tb;
                            lds_2f2(sS + (j_l * 132 + c), sa, sb);
                            if constexpr (kQ) {
                                const ulonglong2 tv = __ldg(reinterpret_cast<const ulonglong2*>(Tg + c));
                                ta = tv.x; tb = tv.y;
                            } else {
                                lds_2f2(sT + (i_l * 128 + c), ta, tb);
                            }
                            const f2 x0 = add2(add2(pk2u(r[k4 * 4 + 0], r[k4 * 4 + 1]), sa), ta);
                            const f2 x1 = add2(add2(pk2u(r[k4 * 4 + 2], r[k4 * 4 + 3]), sb), tb);
                            upk2u(x0, r[k4 * 4 + 0], r[k4 * 4 + 1]);
                            upk2u(x1, r[k4 * 4 + 2], r[k4 * 4 + 3]);
                            s1a = add2(s1a, x0); s1b = add2(s1b, x1);
                            s2a = fma2(x0, x0, s2a); s2b = fma2(x1, x1, s2b);
                        }
#pragma unroll
                        for (int e = 0; e < 16; ++e) xr[hf * 16 + e] = r[e];
                    }
                    sStat[(sx * 4 + q) * 128 + row] = make_float2(hsum2(add2(s1a, s1b)), hsum2(add2(s2a, s2b)));
                };
                if (mode) e1_pass1(std::true_type{});
                else e1_pass1(std::false_type{});
                tc_fence_before();
                TR(2, g);
                row_group_sync(lg);                                // statistics exchanged, D1 reads of these lanes retired
                TR(3, g);
                tc_fence_after();
                {
                    const float2 p0 = sStat[(sx * 4 + 0) * 128 + row], p1 = sStat[(sx * 4 + 1) * 128 + row];
                    const float2 p2 = sStat[(sx * 4 + 2) * 128 + row], p3 = sStat[(sx * 4 + 3) * 128 + row];
                    const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.f / 128.f);
                    const float var = fmaxf(((p0.y + p1.y) + (p2.y + p3.y)) * (1.f / 128.f) - mean * mean, 0.f);
                    const float rstd = rsqrtf(var + kEps);
                    const f2 r2 = pk2(rstd, rstd), n2 = pk2(-mean * rstd, -mean * rstd);
                    uint32_t (&rr)[32] = xr;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t* r = rr + hf * 16;
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const int c = col0 + hf * 16 + k4 * 4;
                            f2 ga, gb, ba, bb;
                            {
                                lds_2f2(sP + (P_MEM_G * 128 + c), ga, gb);
                                lds_2f2(sP + (P_MEM_B * 128 + c), ba, bb);
                            }
                            const f2 y0 = fma2(fma2(pk2u(r[k4 * 4 + 0], r[k4 * 4 + 1]), r2, n2), ga, ba);
                            const f2 y1 = fma2(fma2(pk2u(r[k4 * 4 + 2], r[k4 * 4 + 3]), r2, n2), gb, bb);
                            // hi = ReLU(y) truncated to fp16 (<= y for y >= 0, 0 for y < 0); lo = ReLU(y - hi)
                            const uint32_t h0 = cvt_rz_relu_h2(y0), h1 = cvt_rz_relu_h2(y1);
                            hi[k4 * 2 + 0] = h0;
                            hi[k4 * 2 + 1] = h1;
                            lo[k4 * 2 + 0] = cvt_rn_relu_h2(sub2(y0, h2_to_f2(h0)));
                            lo[k4 * 2 + 1] = cvt_rn_relu_h2(sub2(y1, h2_to_f2(h1)));
                        }
                        // K elements [32q + 16hf, +16) -> cells [16q + 8hf, +8) of the hi block and of the lo block
                        TMEM_ST_X8(tmem + lane_base + d1c + q * 16 + hf * 8, hi);
                        TMEM_ST_X8(tmem + lane_base + d1c + 64 + q * 16 + hf * 8, lo);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                tc_fence_before();
                TR(4, g);
                handoff_arrive(kBarA);                             // this warp's slice of the A operand is in TMEM
                TR(13, g);
                if (ch > ch0) attend(par ^ 1, (i0 - kstep + i_l) < N); // attention epilogue of the previous tile, under the W_pe MMAs
                TR(6, g);
                handoff_arrive(kBarK);                             // Dk/Dv free: the issuer may run the K|V MMAs of this tile

                // ---- epilogue 2a: edge' = LN_e(edge + ReLU(LN_p(Dpe + b_pe))) in place ----
                if (a.has_edge) {
                    mbar_wait(bar_m2a, par, a.err, E_MMA2);
                    TR(7, g);
                    tc_fence_after();
                    // A/B build for the next round: Dpe + b stays in 32 registers from pass A to pass C (no TMEM scratch round
                    // trips, b_pe loaded and added once instead of twice).  Not yet run on hardware.
                    uint32_t er[2][16];
                    {   // pass A: statistics of Dpe + b
                        f2 s1a = 0ull, s1b = 0ull, s2a = 0ull, s2b = 0ull;
                        uint32_t (&rr)[2][16] = er;
                        TMEM_LD_X16_NM(scr_t, rr[0]);
                        TMEM_LD_X16_NM(scr_t + 16, rr[1]);
                        TMEM_WAIT_LD_R16(rr[0]);
                        TMEM_PIN_R16(rr[1]);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            uint32_t (&r)[16] = rr[hf];
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                f2 ba, bb;
                                lds_2f2(sP + (P_BPE * 128 + col0 + hf * 16 + k4 * 4), ba, bb);
                                const f2 x0 = add2(pk2u(r[k4 * 4 + 0], r[k4 * 4 + 1]), ba);
                                const f2 x1 = add2(pk2u(r[k4 * 4 + 2], r[k4 * 4 + 3]), bb);
                                upk2u(x0, r[k4 * 4 + 0], r[k4 * 4 + 1]);
                                upk2u(x1, r[k4 * 4 + 2], r[k4 * 4 + 3]);
                                s1a = add2(s1a, x0); s1b = add2(s1b, x1);
                                s2a = fma2(x0, x0, s2a); s2b = fma2(x1, x1, s2b);
                            }
                        }
                        sStat[(sy * 4 + q) * 128 + row] = make_float2(hsum2(add2(s1a, s1b)), hsum2(add2(s2a, s2b)));
                    }
                    TR(8, g);
                    row_group_sync(lg);
                    TR(9, g);
                    {   // pass B: x = edge + ReLU(LN_p(Dpe + b)) written back to the same cells, statistics of x
                        const float2 p0 = sStat[(sy * 4 + 0) * 128 + row], p1 = sStat[(sy * 4 + 1) * 128 + row];
                        const float2 p2 = sStat[(sy * 4 + 2) * 128 + row], p3 = sStat[(sy * 4 + 3) * 128 + row];
                        const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.f / 128.f);
                        const float var = fmaxf(((p0.y + p1.y) + (p2.y + p3.y)) * (1.f / 128.f) - mean * mean, 0.f);
                        const float rstd = rsqrtf(var + kEps);
                        const f2 r2 = pk2(rstd, rstd), n2 = pk2(-mean * rstd, -mean * rstd);
                        f2 s1a = 0ull, s1b = 0ull, s2a = 0ull, s2b = 0ull;
                        uint32_t (&rr)[2][16] = er;
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            uint32_t (&r)[16] = rr[hf];
#pragma unroll
                            for (int c8 = 0; c8 < 2; ++c8) {
                                const int c = col0 + hf * 16 + c8 * 8;
                                const uint4 eu = ld_shared_v4(tX + tile_off + sw128(row, chunk0 + hf * 2 + c8));
                                const uint32_t ev[4] = {eu.x, eu.y, eu.z, eu.w};
#pragma unroll
                                for (int h4 = 0; h4 < 2; ++h4) {
                                    f2 g0, g1, a0, a1;
                                    lds_2f2(sP + (P_PE_G * 128 + c + h4 * 4), g0, g1);
                                    lds_2f2(sP + (P_PE_B * 128 + c + h4 * 4), a0, a1);
                                    const int k = c8 * 8 + h4 * 4;
                                    float u0, u1, u2, u3;
                                    upk2(fma2(fma2(pk2u(r[k + 0], r[k + 1]), r2, n2), g0, a0), u0, u1);
                                    upk2(fma2(fma2(pk2u(r[k + 2], r[k + 3]), r2, n2), g1, a1), u2, u3);
                                    const f2 x0 = add2(h2_to_f2(ev[h4 * 2 + 0]), pk2(fmaxf(u0, 0.f), fmaxf(u1, 0.f)));
                                    const f2 x1 = add2(h2_to_f2(ev[h4 * 2 + 1]), pk2(fmaxf(u2, 0.f), fmaxf(u3, 0.f)));
                                    upk2u(x0, r[k + 0], r[k + 1]);
                                    upk2u(x1, r[k + 2], r[k + 3]);
                                    s1a = add2(s1a, x0); s1b = add2(s1b, x1);
                                    s2a = fma2(x0, x0, s2a); s2b = fma2(x1, x1, s2b);
                                }
                            }
                        }
                        sStat[(sx * 4 + q) * 128 + row] = make_float2(hsum2(add2(s1a, s1b)), hsum2(add2(s2a, s2b)));
                    }
                    TR(10, g);
                    row_group_sync(lg);
                    TR(11, g);
                    {   // pass C: edge' = LN_e(x) -> fp16, in place in the smem tile
                        const float2 p0 = sStat[(sx * 4 + 0) * 128 + row], p1 = sStat[(sx * 4 + 1) * 128 + row];
                        const float2 p2 = sStat[(sx * 4 + 2) * 128 + row], p3 = sStat[(sx * 4 + 3) * 128 + row];
                        const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.f / 128.f);
                        const float var = fmaxf(((p0.y + p1.y) + (p2.y + p3.y)) * (1.f / 128.f) - mean * mean, 0.f);
                        const float rstd = rsqrtf(var + kEps);
                        const f2 r2 = pk2(rstd, rstd), n2 = pk2(-mean * rstd, -mean * rstd);
                        uint32_t (&rr)[2][16] = er;
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            uint32_t (&r)[16] = rr[hf];
#pragma unroll
                            for (int c8 = 0; c8 < 2; ++c8) {
                                const int c = col0 + hf * 16 + c8 * 8;
                                f2 g0, g1, g2, g3, b0, b1, b2, b3;
                                lds_2f2(sP + (P_NE_G * 128 + c), g0, g1);
                                lds_2f2(sP + (P_NE_G * 128 + c + 4), g2, g3);
                                lds_2f2(sP + (P_NE_B * 128 + c), b0, b1);
                                lds_2f2(sP + (P_NE_B * 128 + c + 4), b2, b3);
                                const uint32_t* x = r + c8 * 8;
                                uint4 u;
                                u.x = cvt_rn_h2(fma2(fma2(pk2u(x[0], x[1]), r2, n2), g0, b0));
                                u.y = cvt_rn_h2(fma2(fma2(pk2u(x[2], x[3]), r2, n2), g1, b1));
                                u.z = cvt_rn_h2(fma2(fma2(pk2u(x[4], x[5]), r2, n2), g2, b2));
                                u.w = cvt_rn_h2(fma2(fma2(pk2u(x[6], x[7]), r2, n2), g3, b3));
                                st_shared_v4(tX + tile_off + sw128(row, chunk0 + hf * 2 + c8), u);
                            }
                        }
                    }
                }

                if (a.has_edge) fence_proxy_async();               // edge' (generic stores) -> visible to the TMA store
                else mbar_wait(bar_m2a, par, a.err, E_MMA2);       // keeps every warp within one tile of the issuer (named barrier kBarE)
                tc_fence_before();
                TR(12, g);
                handoff_arrive(kBarE);                             // tile done: edge' in smem, all TMEM reads retired
            }

            attend((g - 1) & 1, ((ch1 - 1) * kstep + i_l) < N);        // attention epilogue of the work item's last tile
            TR(26, g);
            float acc[32];
#pragma unroll
            for (int k = 0; k < 16; ++k) upk2(acc2[k], acc[2 * k], acc[2 * k + 1]);
            // ---- merge the partial softmax states of the 8 key slots of every (query, head) ----
            // lanes l and l^16 hold key slots 2*lg and 2*lg+1 of the same query; in a single-query item all 32 lanes hold
            // keys of that query (full butterfly: lane 0 ends up with the warp's state, lanes 1..15 are ignored below)
            const int off_min = mode ? 1 : 16;
#pragma unroll 1
            for (int off = 16; off >= off_min; off >>= 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float mo = __shfl_xor_sync(0xffffffffu, mrun[h], off);
                    const float lo = __shfl_xor_sync(0xffffffffu, lrun[h], off);
                    const float mn = fmaxf(mrun[h], mo);
                    const float ca = (mrun[h] == -INFINITY) ? 0.f : __expf(mrun[h] - mn);
                    const float cb = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
                    lrun[h] = lrun[h] * ca + lo * cb;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float ao = __shfl_xor_sync(0xffffffffu, acc[h * 16 + k], off);
                        acc[h * 16 + k] = acc[h * 16 + k] * ca + ao * cb;
                    }
                    mrun[h] = mn;
                }
            }
            // three rounds through the (now dead) S|q tile area: row group r publishes, row group 0 accumulates
            float* scr = sS;                                       // [4 quarter][16 j][36] floats = 9 KB <= S|q area
            for (int r = 1; r < 4; ++r) {
                asm volatile("bar.sync 5, 512;" ::: "memory");
                if (lg == r && lane < 16) {
                    float4* d = reinterpret_cast<float4*>(scr + (q * 16 + lane) * 36);
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[k] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
                    d[8] = make_float4(mrun[0], mrun[1], lrun[0], lrun[1]);
                }
                asm volatile("bar.sync 5, 512;" ::: "memory");
                if (lg == 0 && lane < 16) {
                    const float4* d4 = reinterpret_cast<const float4*>(scr + (q * 16 + lane) * 36);
                    const float4 ml = d4[8];
                    const float mo2[2] = {ml.x, ml.y}, lo2[2] = {ml.z, ml.w};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float mn = fmaxf(mrun[h], mo2[h]);
                        const float ca = (mrun[h] == -INFINITY) ? 0.f : __expf(mrun[h] - mn);
                        const float cb = (mo2[h] == -INFINITY) ? 0.f : __expf(mo2[h] - mn);
                        lrun[h] = lrun[h] * ca + lo2[h] * cb;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const float4 v = d4[h * 4 + k4];
                            acc[h * 16 + k4 * 4 + 0] = acc[h * 16 + k4 * 4 + 0] * ca + v.x * cb;
                            acc[h * 16 + k4 * 4 + 1] = acc[h * 16 + k4 * 4 + 1] * ca + v.y * cb;
                            acc[h * 16 + k4 * 4 + 2] = acc[h * 16 + k4 * 4 + 2] * ca + v.z * cb;
                            acc[h * 16 + k4 * 4 + 3] = acc[h * 16 + k4 * 4 + 3] * ca + v.w * cb;
                        }
                        mrun[h] = mn;
                    }
                }
            }
            // writer of a query's state: lanes 0..15 of row group 0 (query j0 + lane)
            const bool wr = (lg == 0 && lane < 16);
            const int jo = lane;
            if (wk.slot >= 0) {
                if (wr) {                                          // key-split item: park the un-normalised state for k_merge_parts
                    float* d = a.part + ((int64_t)wk.slot * 16 + jo) * 144;
#pragma unroll
                    for (int k = 0; k < 32; ++k) d[col0 + k] = acc[k];
                    d[128 + q * 2] = mrun[0]; d[128 + q * 2 + 1] = mrun[1];
                    d[136 + q * 2] = lrun[0]; d[136 + q * 2 + 1] = lrun[1];
                }
            } else if (wr && j0 + jo < N) {
                const int64_t orow = (tok0 + j0 + jo) * 128 + col0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float inv = 1.f / lrun[h];
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const int c = h * 16 + k4 * 4;
                        const float4 bv = *reinterpret_cast<const float4*>(Pm + P_BV * 128 + col0 + c);
                        const float y0 = acc[c + 0] * inv + bv.x, y1 = acc[c + 1] * inv + bv.y;
                        const float y2 = acc[c + 2] * inv + bv.z, y3 = acc[c + 3] * inv + bv.w;
                        const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        uint2 uh, ul;
                        uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
                        ul.x = pack_h2(y0 - f01.x, y1 - f01.y); ul.y = pack_h2(y2 - f23.x, y3 - f23.y);
                        *reinterpret_cast<uint2*>(a.attn_hi + orow + c) = uh;
                        *reinterpret_cast<uint2*>(a.attn_lo + orow + c) = ul;
                    }
                }
            }
            TR(27, g);
            asm volatile("bar.sync 5, 512;" ::: "memory");         // every warp is done with this item's S|q tiles (and the scratch)
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kIssuerWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// merge of key-split work items: one CTA (128 threads = channels) per (split item, query)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_merge_parts(const TcMerge* __restrict__ jobs, const float* __restrict__ part,
                                                     const float* __restrict__ params, __half* __restrict__ attn_hi,
                                                     __half* __restrict__ attn_lo, int Nmax) {
    const TcMerge jb = jobs[blockIdx.x >> 4];
    const int jl = blockIdx.x & 15, c = threadIdx.x, hd = c >> 4;
    if (jb.j0 + jl >= jb.n) return;
    float m = -INFINITY;
    for (int p = 0; p < jb.nparts; ++p) m = fmaxf(m, part[((int64_t)(jb.slot0 + p) * 16 + jl) * 144 + 128 + hd]);
    float l = 0.f, acc = 0.f;
    for (int p = 0; p < jb.nparts; ++p) {
        const float* d = part + ((int64_t)(jb.slot0 + p) * 16 + jl) * 144;
        const float mp = d[128 + hd];
        const float w = (mp == -INFINITY) ? 0.f : __expf(mp - m);
        l += d[136 + hd] * w;
        acc += d[c] * w;
    }
    const float y = acc / l + params[P_BV * 128 + c];
    const __half hh = __float2half_rn(y);
    const int64_t o = ((int64_t)jb.b * Nmax + jb.j0 + jl) * 128 + c;
    attn_hi[o] = hh;
    attn_lo[o] = __float2half_rn(y - __half2float(hh));
}

// ---------------------------------------------------------------------------------------------
// self test kernel: one 128x128x128 product through the same TMA / descriptor / TMEM helpers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
k_tc_selftest(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap, float* out, int* err) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    volatile uint32_t* sTmem = reinterpret_cast<volatile uint32_t*>(sgen + 65536 + 64);
    const uint32_t bar_l = sbase + 65536, bar_m = bar_l + 8, bar_m2 = bar_l + 16;
    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it (TMEM
    // addresses, barrier ids, role tests) in uniform registers instead of an R2UR in front of every tcgen05.ld / st
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) { mbar_init(bar_l, 1); mbar_init(bar_m, 1); mbar_init(bar_m2, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + 65536 + 64), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const uint32_t tmem2 = tmem;
    if (tid == 0) {
        mbar_expect_tx(bar_l, 65536u);
        for (int kb = 0; kb < 2; ++kb) {
            tma_load_2d(sbase + kb * 16384, &amap, bar_l, kb * 64, 0);
            tma_load_2d(sbase + 32768 + kb * 16384, &wmap, bar_l, kb * 64, 0);
        }
        mbar_wait(bar_l, 0, err, E_LOAD_EDGE);
        tc_fence_after();
        for (int kk = 0; kk < 8; ++kk) {
            const uint32_t ko = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
            umma_f16(tmem, umma_desc_sw128(sbase + ko), umma_desc_sw128(sbase + 32768 + ko), umma_idesc_f16(128), kk > 0);
        }
        umma_commit(bar_m);
    }
    mbar_wait(bar_m, 0, err, E_MMA1);
    tc_fence_after();
    const int row = warp * 32 + lane;
    // also read back the A tile through the software swizzle to validate sw128()
    for (int p = 0; p < 4; ++p) {
        uint32_t r[32];
        TMEM_LD_X32(tmem + ((uint32_t)(warp * 32) << 16) + p * 32, r);
        tmem_wait_ld();
        for (int k = 0; k < 32; ++k) out[row * 128 + p * 32 + k] = __uint_as_float(r[k]);
    }
    uint32_t apk[64];     // this row of A as 64 packed fp16 pairs (K order)
    for (int c8 = 0; c8 < 16; ++c8) {
        const uint4 u = ld_shared_v4(sbase + (c8 >> 3) * 16384 + sw128(row, c8 & 7));
        const float2 e0 = unpack_h2(u.x), e1 = unpack_h2(u.y), e2 = unpack_h2(u.z), e3 = unpack_h2(u.w);
        float* o = out + 128 * 128 + row * 128 + c8 * 8;
        o[0] = e0.x; o[1] = e0.y; o[2] = e1.x; o[3] = e1.y; o[4] = e2.x; o[5] = e2.y; o[6] = e3.x; o[7] = e3.y;
        apk[c8 * 4 + 0] = u.x; apk[c8 * 4 + 1] = u.y; apk[c8 * 4 + 2] = u.z; apk[c8 * 4 + 3] = u.w;
    }
    // second product with the A operand in tensor memory: D2 = A_tmem . W^T -> columns [64,192) after A in [0,64)
    {
        const uint32_t lb = (uint32_t)(warp * 32) << 16;
        tc_fence_before();
        __syncthreads();           // every thread has finished reading D (columns 0..127)
        tc_fence_after();
        TMEM_ST_X32(tmem2 + lb, apk);
        TMEM_ST_X32(tmem2 + lb + 32, (apk + 32));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t ko = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                umma_f16_ts(tmem2 + 64, tmem2 + kk * 8, umma_desc_sw128(sbase + 32768 + ko), umma_idesc_f16(128), kk > 0);
            }
            umma_commit(bar_m2);
        }
        mbar_wait(bar_m2, 0, err, E_MMA2);
        tc_fence_after();
        for (int p = 0; p < 4; ++p) {
            uint32_t r[32];
            TMEM_LD_X32(tmem2 + lb + 64 + p * 32, r);
            tmem_wait_ld();
            for (int k = 0; k < 32; ++k) out[2 * 128 * 128 + row * 128 + p * 32 + k] = __uint_as_float(r[k]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
char g_tc_err[256];
const char* tcfail(const char* what, int code) {
    snprintf(g_tc_err, sizeof g_tc_err, "%s (code %d)", what, code);
    return g_tc_err;
}

// 2-D fp16 [rows][128] K-major operand map, box [64 k][128 rows], 128 B swizzle
const char* make_map_2d(void* out, const void* base, int rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
    cuuint64_t dims[2] = {128, (cuuint64_t)rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? nullptr : tcfail("cuTensorMapEncodeTiled(2d)", (int)r);
}
// 2-D fp32 token tensor stq [rows][384] (S | T | q), box [128 floats][8 rows], no swizzle
const char* make_map_stq(void* out, const void* base, int64_t rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
    cuuint64_t dims[2] = {384, (cuuint64_t)rows};
    cuuint64_t strides[1] = {384 * 4};
    cuuint32_t box[2] = {128, 8};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? nullptr : tcfail("cuTensorMapEncodeTiled(stq)", (int)r);
}
// 4-D fp16 edge stream [B][N(i)][N(j)][128], box [64 c][bj queries][bi keys][1] (tile row = bj * key + query)
const char* make_map_edge(void* out, const void* base, int B, int N, int bj, int bi) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
    cuuint64_t dims[4] = {128, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[3] = {256, (cuuint64_t)N * 256, (cuuint64_t)N * N * 256};
    cuuint32_t box[4] = {64, (cuuint32_t)bj, (cuuint32_t)bi, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? nullptr : tcfail("cuTensorMapEncodeTiled(4d)", (int)r);
}
}  // namespace

void tc_free_forward_state(TcForwardState& f) {
    if (f.d_work) cudaFree(f.d_work);
    if (f.d_merge) cudaFree(f.d_merge);
    if (f.d_part) cudaFree(f.d_part);
    f = TcForwardState{};
}

void tc_free(TcWeights& w) {
    for (auto& l : w.layer) {
        if (l.Wcat) cudaFree(l.Wcat);
        if (l.params) cudaFree(l.params);
        l.Wcat = nullptr; l.params = nullptr;
    }
    tc_free_forward_state(w);
    if (w.h_err) cudaFreeHost((void*)w.h_err);
    w.h_err = nullptr; w.d_err = nullptr; w.packed = false;
}

const char* tc_pack_weights(TcWeights& w, const TcHostLayer (&hl)[6]) {
    for (int l = 0; l < 6; ++l) {
        TcLayerDev& d = w.layer[l];
        std::vector<__half> W((size_t)512 * 128, __float2half(0.f));
        std::vector<float> P((size_t)8 * 128, 0.f);
        const TcHostLayer& h = hl[l];
        for (int o = 0; o < 128; ++o)
            for (int k = 0; k < 128; ++k) {
                W[(size_t)o * 128 + k] = __float2half_rn(h.Wmem[(size_t)o * 384 + k]);
                if (h.Wpe) W[(size_t)(128 + o) * 128 + k] = __float2half_rn(h.Wpe[(size_t)o * 128 + k]);
                W[(size_t)(256 + o) * 128 + k] = __float2half_rn(h.Win[(size_t)(128 + o) * 128 + k]);
                W[(size_t)(384 + o) * 128 + k] = __float2half_rn(h.Win[(size_t)(256 + o) * 128 + k]);
            }
        for (int c = 0; c < 128; ++c) {
            P[tc::P_MEM_G * 128 + c] = h.mem_g[c];
            P[tc::P_MEM_B * 128 + c] = h.mem_b[c];
            if (h.Wpe) {
                P[tc::P_BPE * 128 + c] = h.bpe[c];
                P[tc::P_PE_G * 128 + c] = h.pe_g[c];
                P[tc::P_PE_B * 128 + c] = h.pe_b[c];
                P[tc::P_NE_G * 128 + c] = h.ne_g[c];
                P[tc::P_NE_B * 128 + c] = h.ne_b[c];
            }
            P[tc::P_BV * 128 + c] = h.bin[256 + c];
        }
        d.has_edge = h.Wpe ? 1 : 0;
        if (!d.Wcat && cudaMalloc(&d.Wcat, W.size() * sizeof(__half)) != cudaSuccess) return "cudaMalloc(Wcat) failed";
        if (!d.params && cudaMalloc(&d.params, P.size() * sizeof(float)) != cudaSuccess) return "cudaMalloc(params) failed";
        if (cudaMemcpy(d.Wcat, W.data(), W.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return "copy Wcat failed";
        if (cudaMemcpy(d.params, P.data(), P.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return "copy params failed";
        if (const char* e = make_map_2d(d.wmap, d.Wcat, 512)) return e;
    }
    if (!w.d_err) {
        // mapped host word: the code of a trapped launch stays readable after the context is gone
        if (cudaHostAlloc((void**)&w.h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) return "cudaHostAlloc(err) failed";
        *w.h_err = 0;
        if (cudaHostGetDevicePointer((void**)&w.d_err, (void*)w.h_err, 0) != cudaSuccess) return "cudaHostGetDevicePointer(err) failed";
    }
    w.packed = true;
    return nullptr;
}

// Static schedule of the fused layer kernel.  Work items in scene order; a query block that holds one query only
// (n % 16 == 1: the cls token of a 32 x 128 scene) becomes a single-query item of ceil(n / 128) tiles instead of
// ceil(n / 8) tiles that are 15/16 padding.  Every CTA gets a contiguous run of the tile sequence, equal to within one
// tile: an item that straddles a boundary is split along the key axis and the partial softmax states of its parts are
// merged by k_merge_parts.  work[0 .. grid) are the per-CTA range headers.
static const bool kSingleQueryItems = true;
// Default: every CTA gets one contiguous run of the scene-ordered tile sequence.  MIND_TC_SCHED=deal (development) deals
// whole items round by round first and balances only the remainder; measured slower (single-query items pile up on the
// last CTAs: 9.99 vs 9.84 ms per 5 launches).
static bool deal_rounds() { const char* e = getenv("MIND_TC_SCHED"); return e && !strcmp(e, "deal"); }
void tc_build_schedule(const int* n_tokens, int B, int sm_count, std::vector<TcWork>& work, std::vector<TcMerge>& merges,
                       int& n_slots, int& grid) {
    std::vector<TcWork> items;
    int64_t total = 0;
    for (int b = 0; b < B; ++b) {
        const int n = n_tokens[b];
        for (int j0 = 0; j0 < n; j0 += 16) {
            if (kSingleQueryItems && n - j0 == 1 && n > 16) items.push_back(TcWork{b, j0, n, 0, (n + 127) >> 7, -1, 1, 0});
            else items.push_back(TcWork{b, j0, n, 0, (n + 7) >> 3, -1, 0, 0});
            total += items.back().ch1;
        }
    }
    grid = (int)std::max<int64_t>(1, std::min<int64_t>(total, sm_count));
    merges.clear();
    n_slots = 0;
    // per-CTA lists.  Bulk: whole 16-query items dealt round by round (largest first, serpentine), so that at any time
    // neighbouring CTAs stream neighbouring query blocks of the same scenes.  Remainder (last rounds + single-query
    // items): handed out at tile granularity so that every CTA ends within about one tile of the mean.
    const bool kDealRounds = deal_rounds();
    std::vector<std::vector<TcWork>> lists((size_t)grid);
    std::vector<int64_t> load((size_t)grid, 0);
    std::vector<TcWork> bulk, rest;
    for (const TcWork& t : items) (t.mode ? rest : bulk).push_back(t);
    std::stable_sort(bulk.begin(), bulk.end(), [](const TcWork& x, const TcWork& y) { return x.ch1 > y.ch1; });
    const bool uniform = bulk.empty() || bulk.front().ch1 == bulk.back().ch1;
    int rounds = (int)(bulk.size() / (size_t)grid);
    if (!uniform) rounds = std::max(0, rounds - 1);
    if (!kDealRounds) rounds = 0;
    for (int k = 0; k < rounds; ++k)
        for (int c = 0; c < grid; ++c) {
            const int cc = (k & 1) ? grid - 1 - c : c;
            const TcWork& t = bulk[(size_t)k * grid + c];
            lists[cc].push_back(t);
            load[cc] += t.ch1;
        }
    rest.insert(rest.begin(), bulk.begin() + (size_t)rounds * grid, bulk.end());
    if (!kDealRounds)      // contiguous mode: scene order
        std::stable_sort(rest.begin(), rest.end(), [](const TcWork& x, const TcWork& y) { return x.b != y.b ? x.b < y.b : x.j0 < y.j0; });
    // fill every CTA up to its share of the total
    size_t ri = 0;          // next item of `rest`
    int c0 = 0;             // first chunk of rest[ri] not handed out yet
    for (int c = 0; c < grid && ri < rest.size(); ++c) {
        const int64_t target = total * (c + 1) / grid - total * c / grid;
        const bool last = (c == grid - 1);
        while (ri < rest.size()) {
            const TcWork& t = rest[ri];
            const int64_t need = last ? (int64_t)1 << 40 : target - load[c];
            if (need <= 0) break;
            const int left = t.ch1 - c0;
            if (t.mode || (c0 == 0 && left <= need)) {               // whole item (single-query items are never split)
                lists[c].push_back(t);
                load[c] += left;
                ++ri;
                continue;
            }
            if (c0 == 0) merges.push_back(TcMerge{t.b, t.j0, t.n, n_slots, 0, 0, 0, 0});
            const int take = (int)std::min<int64_t>(left, need);
            TcWork x = t;
            x.ch0 = c0; x.ch1 = c0 + take; x.slot = n_slots++;
            lists[c].push_back(x);
            ++merges.back().nparts;
            load[c] += take;
            c0 += take;
            if (c0 == t.ch1) { c0 = 0; ++ri; }
        }
    }
    work.assign((size_t)grid, TcWork{0, 0, 0, 0, 0, -1, 0, 0});
    for (int c = 0; c < grid; ++c) {
        work[c].b = (int)work.size();
        work.insert(work.end(), lists[c].begin(), lists[c].end());
        work[c].j0 = (int)work.size();
    }
}

const char* tc_prepare(TcWeights& w, const std::vector<SceneDesc>& sd, int B, int Nmax, int min_tokens, __half* edge16, HostStage& stage, cudaStream_t st) {
    if (!w.packed) return "weights not packed";
    std::vector<int> ntok((size_t)B);
    bool any = false;
    for (int b = 0; b < B; ++b) {        // scenes below min_tokens belong to the exact tier: no work items
        const int n = sd[b].n_actor + sd[b].n_lane + 1;
        ntok[b] = n >= min_tokens ? n : 0;
        any = any || ntok[b] > 0;
    }
    w.B = B; w.Nmax = Nmax;
    if (!any) { w.grid = 0; w.n_work = 0; w.n_merge = 0; return nullptr; }
    std::vector<TcWork> work;
    std::vector<TcMerge> merges;
    int n_slots = 0, grid = 1;
    tc_build_schedule(ntok.data(), B, w.sm_count, work, merges, n_slots, grid);
    w.grid = grid;
    if ((int)work.size() > w.work_cap) {
        if (w.d_work) cudaFree(w.d_work);
        if (cudaMalloc(&w.d_work, work.size() * sizeof(TcWork)) != cudaSuccess) return "cudaMalloc(work) failed";
        w.work_cap = (int)work.size();
    }
    if (const char* e = stage.upload(w.d_work, work.data(), work.size() * sizeof(TcWork), st)) return e;
    w.n_work = (int)work.size();
    w.n_merge = (int)merges.size();
    if (w.n_merge > 0) {
        if (w.n_merge > w.merge_cap) {
            if (w.d_merge) cudaFree(w.d_merge);
            if (cudaMalloc(&w.d_merge, merges.size() * sizeof(TcMerge)) != cudaSuccess) return "cudaMalloc(merge) failed";
            w.merge_cap = w.n_merge;
        }
        if (n_slots > w.part_cap) {
            if (w.d_part) cudaFree(w.d_part);
            if (cudaMalloc(&w.d_part, (size_t)n_slots * 16 * 144 * sizeof(float)) != cudaSuccess) return "cudaMalloc(part) failed";
            w.part_cap = n_slots;
        }
        if (const char* e = stage.upload(w.d_merge, merges.data(), merges.size() * sizeof(TcMerge), st)) return e;
    }
    if (w.emap_ptr != edge16 || w.emap_B != B || w.emap_N != Nmax) {
        if (const char* e = make_map_edge(w.emap, edge16, B, Nmax, 16, 8)) return e;
        if (const char* e = make_map_edge(w.emapq, edge16, B, Nmax, 1, 128)) return e;
        w.emap_ptr = edge16; w.emap_B = B; w.emap_N = Nmax;
    }
    w.B = B; w.Nmax = Nmax;
    return nullptr;
}

const char* tc_fusion_layer(TcWeights& w, int layer, const float* stq, __half* attn_hi, __half* attn_lo, int sm_count, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc::k_rela_fusion_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES) != cudaSuccess)
            return "cudaFuncSetAttribute(smem) failed";
        attr = true;
    }
    if (w.grid <= 0) return nullptr;      // no scene takes the fused tier
    tc::LayerArgs a;
    a.work = w.d_work; a.n_work = w.n_work; a.stq = stq; a.params = w.layer[layer].params; a.attn_hi = attn_hi; a.attn_lo = attn_lo;
    a.Nmax = w.Nmax; a.has_edge = w.layer[layer].has_edge; a.err = w.d_err; a.part = w.d_part;
    const int grid = w.grid;
    (void)sm_count;
    if (w.tmap_ptr != stq || w.tmap_rows != (int64_t)w.B * w.Nmax) {
        if (const char* e = make_map_stq(w.tmap, stq, (int64_t)w.B * w.Nmax)) return e;
        w.tmap_ptr = stq; w.tmap_rows = (int64_t)w.B * w.Nmax;
    }
    CUtensorMap em, eq, wm, tm;
    memcpy(&em, w.emap, sizeof em);
    memcpy(&eq, w.emapq, sizeof eq);
    memcpy(&wm, w.layer[layer].wmap, sizeof wm);
    memcpy(&tm, w.tmap, sizeof tm);
    tc::k_rela_fusion_tc<<<grid, tc::kThreads, tc::SMEM_BYTES, st>>>(em, eq, wm, tm, a);
    ++g_launches;
    if (w.n_merge > 0) {
        tc::k_merge_parts<<<w.n_merge * 16, 128, 0, st>>>(w.d_merge, w.d_part, w.layer[layer].params, attn_hi, attn_lo, w.Nmax);
        ++g_launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cudaGetErrorString(e);
    return nullptr;
}

#ifdef MIND_PROGRESS
static volatile int* h_progress = nullptr;
extern "C" int mind_progress_init() {
    if (h_progress) return 0;
    if (cudaHostAlloc((void**)&h_progress, 148 * 17 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) return -1;
    for (int i = 0; i < 148 * 17; ++i) h_progress[i] = -1;
    int* d = nullptr;
    if (cudaHostGetDevicePointer((void**)&d, (void*)h_progress, 0) != cudaSuccess) return -2;
    return cudaMemcpyToSymbol(tc::g_progress, &d, sizeof d) == cudaSuccess ? 0 : -3;
}
extern "C" int mind_progress_read(int* out, int capacity) {
    if (!h_progress || capacity < 148 * 17) return -1;
    for (int i = 0; i < 148 * 17; ++i) out[i] = h_progress[i];
    return 148 * 17;
}
#endif

#ifdef MIND_TRACE
extern "C" int mind_trace_read(long long* host, int capacity) {
    const int n = tc::kTraceTiles * 17 * tc::kTracePts;
    if (capacity < n) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host, tc::g_trace, sizeof(long long) * n) == cudaSuccess ? n : -2;
}
#endif

const char* tc_selftest(const float* A_host, const float* W_host, float* D_host) {
    __half *dA = nullptr, *dW = nullptr;
    float* dO = nullptr;
    int* dErr = nullptr;
    std::vector<__half> hA(128 * 128), hW(128 * 128);
    for (int i = 0; i < 128 * 128; ++i) { hA[i] = __float2half_rn(A_host[i]); hW[i] = __float2half_rn(W_host[i]); }
    if (cudaMalloc(&dA, 32768) != cudaSuccess || cudaMalloc(&dW, 32768) != cudaSuccess ||
        cudaMalloc(&dO, 3 * 65536) != cudaSuccess || cudaMalloc(&dErr, 4) != cudaSuccess)
        return "selftest cudaMalloc failed";
    cudaMemcpy(dA, hA.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemset(dErr, 0, 4);
    cudaMemset(dO, 0, 3 * 65536);
    alignas(64) CUtensorMap am, wm;
    if (const char* e = make_map_2d(&am, dA, 128)) return e;
    if (const char* e = make_map_2d(&wm, dW, 128)) return e;
    const int smem = 65536 + 128 + 1024;
    if (cudaFuncSetAttribute(tc::k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return "selftest attribute failed";
    tc::k_tc_selftest<<<1, 128, smem>>>(am, wm, dO, dErr);
    ++g_launches;
    cudaError_t e = cudaDeviceSynchronize();
    const char* ret = nullptr;
    if (e != cudaSuccess) {
        int code = 0;
        cudaMemcpy(&code, dErr, 4, cudaMemcpyDeviceToHost);
        ret = tcfail(cudaGetErrorString(e), code);
    } else {
        cudaMemcpy(D_host, dO, 3 * 65536, cudaMemcpyDeviceToHost);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dO); cudaFree(dErr);
    return ret;
}

}  // namespace mind
