// Token-side tail of a rela-fusion layer and head of the next one as ONE persistent tcgen05 kernel
// (reference RelaFusionLayer, planners/mind/networks/network.py:178-179, 222-232, and _build_memory's node terms :197-199):
//     x1 = LN2(x + out_proj(attn))                       attn = the fused pair kernel's output, fp16 (hi, lo)
//     x2 = LN3(x1 + linear2(ReLU(linear1(x1))))          FFN 128 -> 256 -> 128
//     [S | T | q] of the NEXT layer = x2 . Wstq^T + b    (384 columns; skipped behind the last layer)
// for a tile of 128 token rows: nothing between the attention output and the next layer's S | T | q rows touches HBM
// except x itself.  Replaces, per layer, 4 GEMM-engine launches + 2 LayerNorm launches (each a round trip of the token
// tensor).  Every contraction is a 3-term fp16 hi/lo product (fp32-equivalent), activation operands live in TMEM as
// (hi, lo) pairs written by the epilogue warps, the 16 weight matrices of a tile stream through a 3-stage ring.
// Same roles and hand-offs as lane_tc.cu: warp 0 issues TMA + tcgen05.mma, warps 1-16 are the epilogue.
#include "tc_gemm.h"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstring>

namespace mind {
namespace node {
using namespace mind::tcp;

constexpr int kThreads = 544;
constexpr uint32_t kStage = 32768;                       // one 128 x 128 fp16 matrix (2 k-blocks of [128 rows][128 B])
constexpr int kRing = 4;                                 // weight stages in flight: the stream out of L2 is latency-bound
constexpr uint32_t SM_RING = 0;
constexpr uint32_t SM_X = kRing * kStage;                    // fp32 [128 rows][128 ch] parked x1, chunks XOR-swizzled by row
constexpr uint32_t SM_P = SM_X + 65536;                  // float [11][128]: bo n2g n2b | b1 (2) | b2 n3g n3b | bstq (3)
constexpr uint32_t SM_STAT = SM_P + 11 * 512;            // float2 [2][4][128]
constexpr uint32_t SM_BAR = SM_STAT + 8192;              // w_full[kRing] w_empty[kRing] d_full
constexpr uint32_t SM_TMEM = SM_BAR + 16 * kRing + 32;     // + d_full, d_stq[2]
constexpr uint32_t SMEM_BYTES = SM_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
constexpr int kBarHand = 6;
// TMEM: A operand (128-wide activation) [0,128) = hi [0,64) | lo [64,128); accumulator [128,256);
// FFN hidden as a 256-wide A operand [256,512) = hi [256,384) | lo [384,512)
constexpr uint32_t TM_A = 0, TM_D = 128, TM_H = 256;
enum { E_WFULL = 32, E_WEMPTY = 33, E_DFULL = 34 };
enum { P_BO = 0, P_N2G, P_N2B, P_B1, P_B1B, P_B2, P_N3G, P_N3B, P_BS, P_BT, P_BQ };

struct Args {
    const __half* ah; const __half* al;   // [rows][128] attention output (hi, lo)
    float* x;                             // [rows][128] token state, updated in place
    float* stq;                           // [rows][384] S | T | q of the next layer (null behind the last layer)
    const float* params;                  // [11][128]
    int64_t rows;
    int n_mats;                           // 16 with the next layer's projection, 10 without
    int* err;
};

__device__ __forceinline__ void hand_arrive() { asm volatile("bar.arrive %0, 544;" ::"r"(kBarHand) : "memory"); }
__device__ __forceinline__ void hand_sync() { asm volatile("bar.sync %0, 544;" ::"r"(kBarHand) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1) k_node_chain_tc(const __grid_constant__ CUtensorMap wmap, Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    float* sP = reinterpret_cast<float*>(sgen + SM_P);
    float2* sStat = reinterpret_cast<float2*>(sgen + SM_STAT);
    volatile uint32_t* sTmem = reinterpret_cast<volatile uint32_t*>(sgen + SM_TMEM);
    const uint32_t bar0 = sbase + SM_BAR, bar_e = bar0 + 8 * kRing, bar_d = bar0 + 16 * kRing, bar_s1 = bar_d + 8, bar_s2 = bar_d + 16;      // w_full[s] = bar0 + 8 s, w_empty[s] = bar_e + 8 s
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kRing; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_e + 8 * s, 1); }
        mbar_init(bar_d, 1); mbar_init(bar_s1, 1); mbar_init(bar_s2, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_TMEM), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 11 * 128; i += kThreads) sP[i] = a.params[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const int n_tiles = (int)((a.rows + 127) / 128);
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const bool with_stq = a.n_mats == 16;

    if (warp == 0) {
        // =============================== issuer warp ===============================
        const uint32_t total = (uint32_t)my_tiles * (uint32_t)a.n_mats;
        uint32_t ld = 0, use = 0;
        const uint32_t id128 = umma_idesc_f16(128);
        auto load_stage = [&](uint32_t s) {
            const uint32_t slot = s % kRing, dst = sbase + SM_RING + slot * kStage, full = bar0 + 8 * slot;
            if (s >= kRing) mbar_wait(bar_e + 8 * slot, ((s / kRing) - 1u) & 1u, a.err, E_WEMPTY);
            mbar_expect_tx(full, kStage);
            const int mrow = (int)(s % (uint32_t)a.n_mats) * 128;
            tma_load_2d(dst, &wmap, full, 0, mrow);
            tma_load_2d(dst + 16384, &wmap, full, 64, mrow);
        };
        auto take_stage = [&]() -> uint32_t {
            while (ld < total && ld < use + kRing) load_stage(ld++);
            mbar_wait(bar0 + 8 * (use % kRing), (use / kRing) & 1u, a.err, E_WFULL);
            tc_fence_after();
            return sbase + SM_RING + (use % kRing) * kStage;
        };
        auto release_stage = [&]() { umma_commit(bar_e + 8 * (use % kRing)); ++use; };
        // 8 k-steps of one 128-wide K block: D (+)= A[acol + 8 kk] . W_stage
        auto mma_k128d = [&](uint32_t dcol, uint32_t acol, uint32_t stage, bool first) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t kw = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                umma_f16_ts(tmem + dcol, tmem + acol + kk * 8, umma_desc_sw128(stage + kw), id128, (first && kk == 0) ? 0u : 1u);
            }
        };
        auto mma_k128 = [&](uint32_t acol, uint32_t stage, bool first) { mma_k128d(TM_D, acol, stage, first); };
        // 3-term product of the 128-wide operand at TM_A with the next (hi, lo) stage pair
        auto round_k128 = [&]() {
            const uint32_t sh = take_stage();
            mma_k128(TM_A, sh, true);
            mma_k128(TM_A + 64, sh, false);
            release_stage();
            const uint32_t sl = take_stage();
            mma_k128(TM_A, sl, false);
            release_stage();
            umma_commit(bar_d);
        };
        if (my_tiles > 0 && elect_one())
            while (ld < total && ld < kRing) load_stage(ld++);
        __syncwarp();
        for (int t = 0; t < my_tiles; ++t) {
            for (int r = 0; r < 3; ++r) {          // out-proj, linear1 columns [0,128), linear1 columns [128,256)
                hand_sync();
                if (elect_one()) { tc_fence_after(); round_k128(); }
                __syncwarp();
            }
            hand_sync();                            // linear2: K = 256 operand at TM_H (hi [0,128) | lo [128,256) cells)
            if (elect_one()) {
                tc_fence_after();
                uint32_t s0 = take_stage();         // W2[:, 0:128] hi
                mma_k128(TM_H, s0, true);
                mma_k128(TM_H + 128, s0, false);
                release_stage();
                s0 = take_stage();                  // W2[:, 128:256] hi
                mma_k128(TM_H + 64, s0, false);
                mma_k128(TM_H + 128 + 64, s0, false);
                release_stage();
                s0 = take_stage();                  // W2[:, 0:128] lo
                mma_k128(TM_H, s0, false);
                release_stage();
                s0 = take_stage();                  // W2[:, 128:256] lo
                mma_k128(TM_H + 64, s0, false);
                release_stage();
                umma_commit(bar_d);
            }
            __syncwarp();
            if (with_stq) {
                // S, T, q blocks of the next layer: one hand-off, three accumulators (the FFN operand columns are free once
                // linear2 has completed), so the epilogue's global stores of one block run under the MMAs of the next
                hand_sync();
                if (elect_one()) {
                    tc_fence_after();
                    for (int blk = 0; blk < 3; ++blk) {
                        const uint32_t dcol = blk == 0 ? TM_D : (blk == 1 ? TM_H : TM_H + 128);
                        const uint32_t sh = take_stage();
                        mma_k128d(dcol, TM_A, sh, true);
                        mma_k128d(dcol, TM_A + 64, sh, false);
                        release_stage();
                        const uint32_t sl = take_stage();
                        mma_k128d(dcol, TM_A, sl, false);
                        release_stage();
                        umma_commit(blk == 0 ? bar_d : (blk == 1 ? bar_s1 : bar_s2));
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        const int ew = warp - 1;
        const int q = ew >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
        const int col0 = q * 32;
        uint32_t rounds = 0, lns = 0;
        float v[32];

        auto wait_d = [&]() {
            mbar_wait(bar_d, rounds & 1u, a.err, E_DFULL);
            ++rounds;
            tc_fence_after();
        };
        auto load_dc = [&](uint32_t dcol, int pb) {
            uint32_t r[32];
            TMEM_LD_X32(tmem + lane_base + dcol + col0, r);
            tmem_wait_ld();
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b = *reinterpret_cast<const float4*>(sP + pb * 128 + col0 + k4 * 4);
                v[k4 * 4 + 0] = __uint_as_float(r[k4 * 4 + 0]) + b.x; v[k4 * 4 + 1] = __uint_as_float(r[k4 * 4 + 1]) + b.y;
                v[k4 * 4 + 2] = __uint_as_float(r[k4 * 4 + 2]) + b.z; v[k4 * 4 + 3] = __uint_as_float(r[k4 * 4 + 3]) + b.w;
            }
        };
        auto load_d = [&](int pb) { load_dc(TM_D, pb); };
        uint32_t stq_tiles = 0;
        auto ln = [&](int pg, int pbeta) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) s += v[k];
            const float ml = s * (1.f / 32.f);
            float m2 = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float d = v[k] - ml; m2 = fmaf(d, d, m2); }
            float2* buf = sStat + (lns & 1u) * 512;
            ++lns;
            buf[q * 128 + row] = make_float2(ml, m2);
            row_group_sync(lg);
            const float2 p0 = buf[row], p1 = buf[128 + row], p2 = buf[256 + row], p3 = buf[384 + row];
            const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * 0.25f;
            const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
            const float var = (((p0.y + p1.y) + (p2.y + p3.y)) + 32.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3))) * (1.f / 128.f);
            const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 g = *reinterpret_cast<const float4*>(sP + pg * 128 + col0 + k4 * 4);
                const float4 b = *reinterpret_cast<const float4*>(sP + pbeta * 128 + col0 + k4 * 4);
                v[k4 * 4 + 0] = (v[k4 * 4 + 0] - mean) * rstd * g.x + b.x; v[k4 * 4 + 1] = (v[k4 * 4 + 1] - mean) * rstd * g.y + b.y;
                v[k4 * 4 + 2] = (v[k4 * 4 + 2] - mean) * rstd * g.z + b.z; v[k4 * 4 + 3] = (v[k4 * 4 + 3] - mean) * rstd * g.w + b.w;
            }
        };
        // v -> fp16 (hi, lo) operand cells [cell0, cell0 + 16) of the halves at `hi_col` / `lo_col`
        auto store_a = [&](uint32_t hi_col, uint32_t lo_col) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                const float2 f = __half22float2(h);
                hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                lo[j] = pack_h2(v[2 * j] - f.x, v[2 * j + 1] - f.y);
            }
            TMEM_ST_X16(tmem + lane_base + hi_col, hi);
            TMEM_ST_X16(tmem + lane_base + lo_col, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };

        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t grow = (int64_t)tile * 128 + row;
            const bool valid = grow < a.rows;
            // ---- attention output of this row (already an fp16 hi / lo pair) -> A operand ----
            {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { hi[j] = 0u; lo[j] = 0u; }
                if (valid) {
                    const uint4* ph = reinterpret_cast<const uint4*>(a.ah + grow * 128 + col0);
                    const uint4* pl = reinterpret_cast<const uint4*>(a.al + grow * 128 + col0);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 h = __ldg(ph + c), l = __ldg(pl + c);
                        hi[c * 4] = h.x; hi[c * 4 + 1] = h.y; hi[c * 4 + 2] = h.z; hi[c * 4 + 3] = h.w;
                        lo[c * 4] = l.x; lo[c * 4 + 1] = l.y; lo[c * 4 + 2] = l.z; lo[c * 4 + 3] = l.w;
                    }
                }
                TMEM_ST_X16(tmem + lane_base + TM_A + q * 16, hi);
                TMEM_ST_X16(tmem + lane_base + TM_A + 64 + q * 16, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            hand_arrive();
            // ---- x1 = LN2(x + out_proj(attn) + b) : parked in shared memory, A operand of linear1 ----
            wait_d();
            load_d(P_BO);
            if (valid) {
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 xv = *reinterpret_cast<const float4*>(a.x + grow * 128 + col0 + k4 * 4);
                    v[k4 * 4] += xv.x; v[k4 * 4 + 1] += xv.y; v[k4 * 4 + 2] += xv.z; v[k4 * 4 + 3] += xv.w;
                }
            }
            ln(P_N2G, P_N2B);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const int ch = (q * 8 + k4) ^ (row & 7);
                *reinterpret_cast<float4*>(sgen + SM_X + row * 512 + ch * 16) = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
            }
            store_a(TM_A + q * 16, TM_A + 64 + q * 16);
            tc_fence_before();
            hand_arrive();
            // ---- FFN hidden: ReLU(linear1(x1) + b), two 128-column halves -> 256-wide A operand ----
            for (int half = 0; half < 2; ++half) {
                wait_d();
                load_d(P_B1 + half);
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
                store_a(TM_H + half * 64 + q * 16, TM_H + 128 + half * 64 + q * 16);
                tc_fence_before();
                hand_arrive();
            }
            // ---- x2 = LN3(x1 + linear2(hidden) + b) -> x (global) and, in front of another layer, the A operand ----
            wait_d();
            load_d(P_B2);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const int ch = (q * 8 + k4) ^ (row & 7);
                const float4 xv = *reinterpret_cast<const float4*>(sgen + SM_X + row * 512 + ch * 16);
                v[k4 * 4] += xv.x; v[k4 * 4 + 1] += xv.y; v[k4 * 4 + 2] += xv.z; v[k4 * 4 + 3] += xv.w;
            }
            ln(P_N3G, P_N3B);
            if (valid) {
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4)
                    *reinterpret_cast<float4*>(a.x + grow * 128 + col0 + k4 * 4) = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
            }
            if (with_stq) {
                store_a(TM_A + q * 16, TM_A + 64 + q * 16);
                tc_fence_before();
                hand_arrive();
                for (int blk = 0; blk < 3; ++blk) {          // S | T(+b_mem) | q/4 of the next layer
                    if (blk == 0) {
                        wait_d();
                    } else {
                        mbar_wait(blk == 1 ? bar_s1 : bar_s2, stq_tiles & 1u, a.err, E_DFULL);
                        tc_fence_after();
                    }
                    load_dc(blk == 0 ? TM_D : (blk == 1 ? TM_H : TM_H + 128), P_BS + blk);
                    if (valid) {
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4)
                            *reinterpret_cast<float4*>(a.stq + grow * 384 + blk * 128 + col0 + k4 * 4) =
                                make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
                    }
                }
                ++stq_tiles;
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace node

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
void node_chain_free(NodeChain& n) {
    for (auto& l : n.layer) {
        if (l.W) cudaFree(l.W);
        if (l.params) cudaFree(l.params);
        l.W = nullptr; l.params = nullptr;
    }
    if (n.d_err) cudaFree(n.d_err);
    n.d_err = nullptr; n.ready = false;
}

// host fp32 weights of one layer (and the fused [S | T | q/4] projection of the NEXT layer) -> streamed (hi, lo) matrices
const char* node_chain_pack(NodeChain& n, int l, const float* Wo, const float* bo, const float* n2g, const float* n2b, const float* W1,
                            const float* b1, const float* W2, const float* b2, const float* n3g, const float* n3b,
                            const float* Wstq_next, const float* bstq_next) {
    NodeChainLayer& d = n.layer[l];
    d.n_mats = Wstq_next ? 16 : 10;
    std::vector<__half> W((size_t)d.n_mats * 128 * 128, __float2half(0.f));
    std::vector<float> P((size_t)11 * 128, 0.f);
    // 128 x 128 block of `src` (row stride ld, rows r0.., columns c0..) -> matrix `hi_mat` (hi part) and `lo_mat` (lo part)
    auto put = [&](int hi_mat, int lo_mat, const float* src, int ld, int r0, int c0) {
        for (int o = 0; o < 128; ++o)
            for (int k = 0; k < 128; ++k) {
                const float w = src[(size_t)(r0 + o) * ld + c0 + k];
                const __half h = __float2half_rn(w);
                W[((size_t)hi_mat * 128 + o) * 128 + k] = h;
                W[((size_t)lo_mat * 128 + o) * 128 + k] = __float2half_rn(w - __half2float(h));
            }
    };
    put(0, 1, Wo, 128, 0, 0);
    put(2, 3, W1, 128, 0, 0);              // linear1 rows [0,128)
    put(4, 5, W1, 128, 128, 0);            // linear1 rows [128,256)
    put(6, 8, W2, 256, 0, 0);              // linear2 K columns [0,128): hi -> 6, lo -> 8
    put(7, 9, W2, 256, 0, 128);            // linear2 K columns [128,256): hi -> 7, lo -> 9
    if (Wstq_next)
        for (int b = 0; b < 3; ++b) put(10 + 2 * b, 11 + 2 * b, Wstq_next, 128, 128 * b, 0);
    auto vec = [&](int idx, const float* s, int cnt) { std::memcpy(P.data() + (size_t)idx * 128, s, (size_t)cnt * 4); };
    vec(node::P_BO, bo, 128); vec(node::P_N2G, n2g, 128); vec(node::P_N2B, n2b, 128); vec(node::P_B1, b1, 256);
    vec(node::P_B2, b2, 128); vec(node::P_N3G, n3g, 128); vec(node::P_N3B, n3b, 128);
    if (bstq_next) vec(node::P_BS, bstq_next, 384);
    if (!d.W && cudaMalloc(&d.W, (size_t)16 * 128 * 128 * sizeof(__half)) != cudaSuccess) return "node_chain_pack: cudaMalloc(W) failed";
    if (!d.params && cudaMalloc(&d.params, P.size() * sizeof(float)) != cudaSuccess) return "node_chain_pack: cudaMalloc(params) failed";
    cudaMemcpy(d.W, W.data(), W.size() * sizeof(__half), cudaMemcpyHostToDevice);
    cudaMemcpy(d.params, P.data(), P.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (const char* e = tcg_encode_w(d.wmap, d.W, 128, (int64_t)d.n_mats * 128, 128)) return e;
    if (!n.d_err) {
        if (cudaMalloc(&n.d_err, sizeof(int)) != cudaSuccess) return "node_chain_pack: cudaMalloc(err) failed";
        cudaMemset(n.d_err, 0, sizeof(int));
    }
    if (l == 5) n.ready = true;
    return nullptr;
}

const char* node_chain_run(NodeChain& n, int l, const __half* ah, const __half* al, float* x, float* stq_next, int64_t rows,
                           int sm_count, cudaStream_t st) {
    if (!n.ready) return "node_chain_run: weights not packed";
    if (rows <= 0) return nullptr;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(node::k_node_chain_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)node::SMEM_BYTES) != cudaSuccess)
            return "cudaFuncSetAttribute(node_chain_tc) failed";
        attr = true;
    }
    const NodeChainLayer& d = n.layer[l];
    if ((d.n_mats == 16) != (stq_next != nullptr)) return "node_chain_run: next-layer projection / output mismatch";
    node::Args a;
    a.ah = ah; a.al = al; a.x = x; a.stq = stq_next; a.params = d.params; a.rows = rows; a.n_mats = d.n_mats; a.err = n.d_err;
    CUtensorMap wm;
    memcpy(&wm, d.wmap, sizeof wm);
    const int tiles = (int)((rows + 127) / 128);
    node::k_node_chain_tc<<<std::min(tiles, sm_count), node::kThreads, node::SMEM_BYTES, st>>>(wm, a);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace mind
