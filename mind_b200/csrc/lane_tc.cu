// LaneNet (reference planners/mind/networks/network.py:64-121) as ONE persistent tcgen05 kernel: the whole chain
//   proj (16 -> 128, LN, ReLU) ; 2 x PointAggregateBlock { fc1 (2 x Linear-LN-ReLU) ; max over the polyline's 10 nodes ;
//   fc2 on [h | max] (Linear 256 -> 128, LN, ReLU ; Linear, LN, ReLU) ; LN(x + y) } ; max over nodes
// runs on chip for a tile of 12 polylines (120 node rows + 8 idle TMEM lanes): activations never touch HBM.
//
// * every contraction is a 3-term fp16 hi/lo product (hi.hi + lo.hi + hi.lo, fp32 accumulation in TMEM): fp32-equivalent,
//   like the un-fused path it replaces (run_lane_net_tc: 8 GEMM launches + 9 LayerNorm launches + 2 max launches,
//   every layer a round trip of 330k x 128 activations through HBM);
// * the activation operand of the next contraction is written by the epilogue warps straight into TMEM as an fp16 (hi, lo)
//   pair (A-from-TMEM MMAs), the block input x is parked in TMEM columns as fp32 for the residual;
// * weights (20 matrices of 32 KB per tile) stream out of L2 through a 3-stage shared-memory ring (TMA, 128 B swizzle);
// * a lane quadrant (32 TMEM lanes) holds 3 whole polylines (rows 0-29), so the node-max never crosses a warp's rows; it
//   goes through a swizzled shared-memory tile (one 16-byte chunk per thread and node row).
// Roles: warp 0 = issuer (TMA, every tcgen05.mma), warps 1-16 = epilogue (warp w: TMEM lanes 32*(w&3).., channel quarter
// (w-1)>>2).  One tile is in flight: MMAs and epilogues alternate, hand-offs as in the fused rela-fusion kernel
// (tcgen05.commit -> mbarrier towards the epilogue, named barrier towards the issuer).
#include "tc_gemm.h"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstring>

namespace mind {
namespace lane {
using namespace mind::tcp;

constexpr int kThreads = 544;
constexpr int kPolyPerTile = 12;
constexpr uint32_t kStage = 32768;                       // one 128 x 128 fp16 matrix: 2 k-blocks of [128 rows][128 B]
constexpr int kRing = 4;                                 // weight stages in flight: the stream out of L2 is latency-bound
constexpr uint32_t SM_RING = 0;
constexpr uint32_t SM_H = kRing * kStage;                // fp32 [128 rows][128 ch], 16-byte chunks XOR-swizzled by row
constexpr uint32_t SM_P = SM_H + 65536;                  // float [31][128] biases / LN parameters
constexpr uint32_t SM_M = SM_P + 31 * 512;               // float [12][128] per-polyline max
constexpr uint32_t SM_STAT = SM_M + 12 * 512;            // float2 [2 buffers][4 quarters][128 rows]
constexpr uint32_t SM_BAR = SM_STAT + 8192;              // w_full[kRing] w_empty[kRing] d_full
constexpr uint32_t SM_TMEM = SM_BAR + 16 * kRing + 16;
constexpr uint32_t SM_TOTAL = SM_TMEM + 16;
constexpr uint32_t SMEM_BYTES = SM_TOTAL + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
constexpr int kMatsPerTile = 21;                         // weight stages per tile: proj (k-block 0 of hi | lo in one stage), then 10 per block (hi, lo)
constexpr int kBarHand = 6;                              // epilogue -> issuer hand-off (named barrier, 544 threads)
// TMEM columns: A operand [0,128) (hi [0,64) | lo [64,128)), accumulator [128,256), parked block input x [256,384) fp32,
// second A operand (per-polyline max, broadcast to its rows) [384,512)
constexpr uint32_t TM_A = 0, TM_D = 128, TM_X = 256, TM_A2 = 384;
enum { E_WPROJ = 21, E_WFULL = 22, E_WEMPTY = 23, E_DFULL = 24 };

struct Args {
    const float* lanes;      // [Lp * 10][16] node features (lane polylines followed by the target polylines)
    const float* params;     // [31][128]
    float* out;              // [Lp][128]
    int Lp;
    int* err;
};

__device__ __forceinline__ void hand_arrive() { asm volatile("bar.arrive %0, 544;" ::"r"(kBarHand) : "memory"); }
__device__ __forceinline__ void hand_sync() { asm volatile("bar.sync %0, 544;" ::"r"(kBarHand) : "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 5, 512;" ::: "memory"); }

#define LANE_TMEM_ST_X32(taddr, r) TMEM_ST_X32(taddr, r)

__global__ void __launch_bounds__(kThreads, 1) k_lane_net_tc(const __grid_constant__ CUtensorMap wmap, Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    float* sP = reinterpret_cast<float*>(sgen + SM_P);
    float* sM = reinterpret_cast<float*>(sgen + SM_M);
    float2* sStat = reinterpret_cast<float2*>(sgen + SM_STAT);
    volatile uint32_t* sTmem = reinterpret_cast<volatile uint32_t*>(sgen + SM_TMEM);
    const uint32_t bar0 = sbase + SM_BAR;
    const uint32_t bar_e = bar0 + 8 * kRing, bar_d = bar0 + 16 * kRing;          // w_full[s] = bar0 + 8 s, w_empty[s] = bar_e + 8 s
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kRing; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_e + 8 * s, 1); }
        mbar_init(bar_d, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_TMEM), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 31 * 128; i += kThreads) sP[i] = a.params[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const int n_tiles = (a.Lp + kPolyPerTile - 1) / kPolyPerTile;
    const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        // =============================== issuer warp ===============================
        const uint32_t total = (uint32_t)my_tiles * kMatsPerTile;
        uint32_t ld = 0, use = 0;                     // weight stages loaded / consumed so far (elected lane only)
        const uint32_t id128 = umma_idesc_f16(128);
        auto load_stage = [&](uint32_t s) {
            const uint32_t slot = s % kRing, dst = sbase + SM_RING + slot * kStage, full = bar0 + 8 * slot;
            if (s >= kRing) mbar_wait(bar_e + 8 * slot, ((s / kRing) - 1u) & 1u, a.err, E_WEMPTY);
            mbar_expect_tx(full, kStage);
            const int m = (int)(s % kMatsPerTile);
            if (m == 0) {          // proj: K = 16 lives in k-block 0 of matrices 20 (hi) and 21 (lo)
                tma_load_2d(dst, &wmap, full, 0, 20 * 128);
                tma_load_2d(dst + 16384, &wmap, full, 0, 21 * 128);
            } else {
                tma_load_2d(dst, &wmap, full, 0, (m - 1) * 128);
                tma_load_2d(dst + 16384, &wmap, full, 64, (m - 1) * 128);
            }
        };
        auto take_stage = [&]() -> uint32_t {         // shared-memory address of the next weight stage, loads kept 2 ahead
            while (ld < total && ld < use + kRing) load_stage(ld++);
            mbar_wait(bar0 + 8 * (use % kRing), (use / kRing) & 1u, a.err, E_WFULL);
            tc_fence_after();
            return sbase + SM_RING + (use % kRing) * kStage;
        };
        if (my_tiles > 0 && elect_one())
            while (ld < total && ld < kRing) load_stage(ld++);
        __syncwarp();
        for (int t = 0; t < my_tiles; ++t) {
            // ---- proj: K = 16, one k-step, 3 terms ----
            hand_sync();
            if (elect_one()) {
                tc_fence_after();
                const uint32_t sp = take_stage();
                const uint64_t bh = umma_desc_sw128(sp), bl = umma_desc_sw128(sp + 16384);
                umma_f16_ts(tmem + TM_D, tmem + TM_A, bh, id128, 0);
                umma_f16_ts(tmem + TM_D, tmem + TM_A + 64, bh, id128, 1);
                umma_f16_ts(tmem + TM_D, tmem + TM_A, bl, id128, 1);
                umma_commit(bar_e + 8 * (use % kRing));
                ++use;
                umma_commit(bar_d);
            }
            __syncwarp();
            // ---- 2 blocks x {fc1.0, fc1.3, fc2.0 on [h | max], fc2.3} ----
            for (int r = 0; r < 8; ++r) {
                hand_sync();
                if (elect_one()) {
                    tc_fence_after();
                    const int pairs = ((r & 3) == 2) ? 2 : 1;
                    for (int pp = 0; pp < pairs; ++pp) {
                        const uint32_t ab = tmem + (pp ? TM_A2 : TM_A);
                        const uint32_t sh = take_stage();
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t kw = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                            umma_f16_ts(tmem + TM_D, ab + kk * 8, umma_desc_sw128(sh + kw), id128, (pp | kk) != 0);
                        }
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t kw = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                            umma_f16_ts(tmem + TM_D, ab + 64 + kk * 8, umma_desc_sw128(sh + kw), id128, 1);
                        }
                        umma_commit(bar_e + 8 * (use % kRing));
                        ++use;
                        const uint32_t sl = take_stage();
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t kw = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                            umma_f16_ts(tmem + TM_D, ab + kk * 8, umma_desc_sw128(sl + kw), id128, 1);
                        }
                        umma_commit(bar_e + 8 * (use % kRing));
                        ++use;
                    }
                    umma_commit(bar_d);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        const int ew = warp - 1;
        const int q = ew >> 2;                        // channel quarter [32q, 32q + 32)
        const int lg = warp & 3;                      // TMEM lane quadrant (fixed by the hardware warp id)
        const int row = lg * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
        const int col0 = q * 32;
        const int et = ew * 32 + lane;                // epilogue thread index 0..511
        uint32_t rounds = 0, lns = 0;
        float v[32];

        auto wait_d = [&]() {
            mbar_wait(bar_d, rounds & 1u, a.err, E_DFULL);
            ++rounds;
            tc_fence_after();
        };
        // v = accumulator columns of this thread + bias vector `pb`
        auto load_d = [&](int pb) {
            uint32_t r[32];
            TMEM_LD_X32(tmem + lane_base + TM_D + col0, r);
            tmem_wait_ld();
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b = *reinterpret_cast<const float4*>(sP + pb * 128 + col0 + k4 * 4);
                v[k4 * 4 + 0] = __uint_as_float(r[k4 * 4 + 0]) + b.x; v[k4 * 4 + 1] = __uint_as_float(r[k4 * 4 + 1]) + b.y;
                v[k4 * 4 + 2] = __uint_as_float(r[k4 * 4 + 2]) + b.z; v[k4 * 4 + 3] = __uint_as_float(r[k4 * 4 + 3]) + b.w;
            }
        };
        // LayerNorm over the 128 channels of a row held by the 4 warps of a lane quadrant: local mean and centred sum of
        // squares per 32-channel quarter, merged exactly (Chan) after one exchange through shared memory
        auto ln = [&](int pg, int pbeta, bool relu) {
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
                float y0 = (v[k4 * 4 + 0] - mean) * rstd * g.x + b.x, y1 = (v[k4 * 4 + 1] - mean) * rstd * g.y + b.y;
                float y2 = (v[k4 * 4 + 2] - mean) * rstd * g.z + b.z, y3 = (v[k4 * 4 + 3] - mean) * rstd * g.w + b.w;
                if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                v[k4 * 4 + 0] = y0; v[k4 * 4 + 1] = y1; v[k4 * 4 + 2] = y2; v[k4 * 4 + 3] = y3;
            }
        };
        // v -> fp16 (hi, lo) A operand at TMEM columns `ac`: K elements [32q, 32q+32) -> cells [16q, 16q+16) of each half
        auto store_a = [&](uint32_t ac, const float* x) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                const float2 f = __half22float2(h);
                hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                lo[j] = pack_h2(x[2 * j] - f.x, x[2 * j + 1] - f.y);
            }
            TMEM_ST_X16(tmem + lane_base + ac + q * 16, hi);
            TMEM_ST_X16(tmem + lane_base + ac + 64 + q * 16, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        // v -> shared-memory tile (fp32, chunk index XOR row & 7: 16-byte stores of 32 rows spread over all banks)
        auto stage_h = [&]() {
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const int ch = (q * 8 + k4) ^ (row & 7);
                *reinterpret_cast<float4*>(sgen + SM_H + row * 512 + ch * 16) = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
            }
        };
        // max over the 10 node rows of polyline `p` (tile-local 0..11) for channels [4 c4, 4 c4 + 4)
        auto poly_max = [&](int p, int c4) -> float4 {
            const int r0 = 32 * (p / 3) + 10 * (p % 3);
            float4 m = *reinterpret_cast<const float4*>(sgen + SM_H + r0 * 512 + ((c4 ^ (r0 & 7)) * 16));
#pragma unroll
            for (int n = 1; n < 10; ++n) {
                const int rr = r0 + n;
                const float4 x = *reinterpret_cast<const float4*>(sgen + SM_H + rr * 512 + ((c4 ^ (rr & 7)) * 16));
                m.x = fmaxf(m.x, x.x); m.y = fmaxf(m.y, x.y); m.z = fmaxf(m.z, x.z); m.w = fmaxf(m.w, x.w);
            }
            return m;
        };

        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // ---- node features of this row -> A operand of the proj product (K = 16: cells [0,8) of each half) ----
            const int pl = lg * 3 + lane / 10;                              // tile-local polyline of this row (lanes 30, 31: none)
            const int64_t pg = (int64_t)tile * kPolyPerTile + pl;
            if (q == 0) {
                float in[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) in[k] = 0.f;
                if (lane < 30 && pg < a.Lp) {
                    const float4* src = reinterpret_cast<const float4*>(a.lanes + (pg * 10 + lane % 10) * 16);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 x = __ldg(src + k4);
                        in[k4 * 4] = x.x; in[k4 * 4 + 1] = x.y; in[k4 * 4 + 2] = x.z; in[k4 * 4 + 3] = x.w;
                    }
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __half2 h = __floats2half2_rn(in[2 * j], in[2 * j + 1]);
                    const float2 f = __half22float2(h);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[j] = pack_h2(in[2 * j] - f.x, in[2 * j + 1] - f.y);
                }
                TMEM_ST_X8(tmem + lane_base + TM_A, hi);
                TMEM_ST_X8(tmem + lane_base + TM_A + 64, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            hand_arrive();
            // ---- proj epilogue: x = ReLU(LN(D + b)) -> parked (fp32) + A operand ----
            wait_d();
            load_d(0);
            ln(1, 2, true);
            {
                uint32_t xr[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) xr[k] = __float_as_uint(v[k]);
                LANE_TMEM_ST_X32(tmem + lane_base + TM_X + col0, xr);
            }
            store_a(TM_A, v);
            tc_fence_before();
            hand_arrive();
            for (int blk = 0; blk < 2; ++blk) {
                const int pb = 3 + 14 * blk;
                // fc1.0
                wait_d();
                load_d(pb + 0);
                ln(pb + 1, pb + 2, true);
                store_a(TM_A, v);
                tc_fence_before();
                hand_arrive();
                // fc1.3 -> h ; per-polyline max of h -> second A operand
                wait_d();
                load_d(pb + 3);
                ln(pb + 4, pb + 5, true);
                store_a(TM_A, v);
                stage_h();
                epi_sync();
                if (et < kPolyPerTile * 32) {
                    const int p = et >> 5, c4 = et & 31;
                    *reinterpret_cast<float4*>(sM + p * 128 + c4 * 4) = poly_max(p, c4);
                }
                epi_sync();
                {
                    float mx[32];
                    const int pm = lane < 30 ? pl : lg * 3;             // idle lanes: any finite row
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const float4 x = *reinterpret_cast<const float4*>(sM + pm * 128 + col0 + k4 * 4);
                        mx[k4 * 4] = x.x; mx[k4 * 4 + 1] = x.y; mx[k4 * 4 + 2] = x.z; mx[k4 * 4 + 3] = x.w;
                    }
                    store_a(TM_A2, mx);
                }
                tc_fence_before();
                hand_arrive();
                // fc2.0 on [h | max]
                wait_d();
                load_d(pb + 6);
                ln(pb + 7, pb + 8, true);
                store_a(TM_A, v);
                tc_fence_before();
                hand_arrive();
                // fc2.3 ; x = LN(x + y)
                wait_d();
                load_d(pb + 9);
                ln(pb + 10, pb + 11, true);
                {
                    uint32_t xr[32];
                    TMEM_LD_X32(tmem + lane_base + TM_X + col0, xr);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] += __uint_as_float(xr[k]);
                }
                ln(pb + 12, pb + 13, false);
                if (blk == 0) {
                    uint32_t xr[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) xr[k] = __float_as_uint(v[k]);
                    LANE_TMEM_ST_X32(tmem + lane_base + TM_X + col0, xr);
                    store_a(TM_A, v);
                    tc_fence_before();
                    hand_arrive();
                } else {
                    // final max over the nodes of each polyline -> lane feature rows
                    stage_h();
                    epi_sync();
                    if (et < kPolyPerTile * 32) {
                        const int p = et >> 5, c4 = et & 31;
                        const int64_t po = (int64_t)tile * kPolyPerTile + p;
                        if (po < a.Lp) *reinterpret_cast<float4*>(a.out + po * 128 + c4 * 4) = poly_max(p, c4);
                    }
                    tc_fence_before();
                    epi_sync();          // the staging tile is free again before the next tile's rows arrive in it
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace lane

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
void lane_tc_free(LaneTc& l) {
    if (l.W) cudaFree(l.W);
    if (l.params) cudaFree(l.params);
    if (l.d_err) cudaFree(l.d_err);
    l = LaneTc{};
}

// host fp32 weights by reference key -> streamed fp16 (hi, lo) matrices + parameter table
const char* lane_tc_pack(LaneTc& l, const std::map<std::string, std::vector<float>>& host) {
    lane_tc_free(l);
    auto get = [&](const std::string& k, size_t n) -> const float* {
        auto it = host.find(k);
        return (it == host.end() || it->second.size() != n) ? nullptr : it->second.data();
    };
    std::vector<__half> W((size_t)22 * 128 * 128, __float2half(0.f));
    std::vector<float> P((size_t)31 * 128, 0.f);
    auto put = [&](int mat, const float* src, int ld, int col_off, int K) {       // rows o, K columns -> mat (hi) and mat + 1 (lo)
        for (int o = 0; o < 128; ++o)
            for (int k = 0; k < K; ++k) {
                const float w = src[(size_t)o * ld + col_off + k];
                const __half h = __float2half_rn(w);
                W[((size_t)mat * 128 + o) * 128 + k] = h;
                W[((size_t)(mat + 1) * 128 + o) * 128 + k] = __float2half_rn(w - __half2float(h));
            }
    };
    auto vec = [&](int idx, const std::string& k) -> bool {
        const float* s = get(k, 128);
        if (!s) return false;
        std::memcpy(P.data() + (size_t)idx * 128, s, 512);
        return true;
    };
    const float* wp = get("lane_net.proj.0.weight", 128 * 16);
    if (!wp) return "lane_tc_pack: missing lane_net.proj.0.weight";
    put(20, wp, 16, 0, 16);
    if (!vec(0, "lane_net.proj.0.bias") || !vec(1, "lane_net.proj.1.weight") || !vec(2, "lane_net.proj.1.bias")) return "lane_tc_pack: missing proj parameters";
    for (int blk = 0; blk < 2; ++blk) {
        const std::string Pn = "lane_net.aggre" + std::to_string(blk + 1) + ".";
        const float* w10 = get(Pn + "fc1.0.weight", 128 * 128);
        const float* w13 = get(Pn + "fc1.3.weight", 128 * 128);
        const float* w20 = get(Pn + "fc2.0.weight", 128 * 256);
        const float* w23 = get(Pn + "fc2.3.weight", 128 * 128);
        if (!w10 || !w13 || !w20 || !w23) return "lane_tc_pack: missing aggregate-block weight";
        const int m0 = 10 * blk, pb = 3 + 14 * blk;
        put(m0 + 0, w10, 128, 0, 128);
        put(m0 + 2, w13, 128, 0, 128);
        put(m0 + 4, w20, 256, 0, 128);        // fc2.0 on h
        put(m0 + 6, w20, 256, 128, 128);      // fc2.0 on the broadcast max
        put(m0 + 8, w23, 128, 0, 128);
        const char* names[14] = {"fc1.0.bias", "fc1.1.weight", "fc1.1.bias", "fc1.3.bias", "fc1.4.weight", "fc1.4.bias", "fc2.0.bias",
                                 "fc2.1.weight", "fc2.1.bias", "fc2.3.bias", "fc2.4.weight", "fc2.4.bias", "norm.weight", "norm.bias"};
        for (int i = 0; i < 14; ++i)
            if (!vec(pb + i, Pn + names[i])) return "lane_tc_pack: missing aggregate-block parameter";
    }
    if (cudaMalloc(&l.W, W.size() * sizeof(__half)) != cudaSuccess) return "lane_tc_pack: cudaMalloc(W) failed";
    if (cudaMalloc(&l.params, P.size() * sizeof(float)) != cudaSuccess) return "lane_tc_pack: cudaMalloc(params) failed";
    if (cudaMalloc(&l.d_err, sizeof(int)) != cudaSuccess) return "lane_tc_pack: cudaMalloc(err) failed";
    cudaMemcpy(l.W, W.data(), W.size() * sizeof(__half), cudaMemcpyHostToDevice);
    cudaMemcpy(l.params, P.data(), P.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(l.d_err, 0, sizeof(int));
    if (const char* e = tcg_encode_w(l.wmap, l.W, 128, 22 * 128, 128)) return e;
    l.ready = true;
    return nullptr;
}

const char* lane_tc_run(LaneTc& l, const float* lanes, int Lp, float* out, int sm_count, cudaStream_t st) {
    if (!l.ready) return "lane_tc_run: weights not packed";
    if (Lp <= 0) return nullptr;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(lane::k_lane_net_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lane::SMEM_BYTES) != cudaSuccess)
            return "cudaFuncSetAttribute(lane_net_tc) failed";
        attr = true;
    }
    lane::Args a;
    a.lanes = lanes; a.params = l.params; a.out = out; a.Lp = Lp; a.err = l.d_err;
    CUtensorMap wm;
    memcpy(&wm, l.wmap, sizeof wm);
    const int tiles = (Lp + lane::kPolyPerTile - 1) / lane::kPolyPerTile;
    lane::k_lane_net_tc<<<std::min(tiles, sm_count), lane::kThreads, lane::SMEM_BYTES, st>>>(wm, a);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace mind
