// Persistent warp-specialised tcgen05 GEMM for sm_100a:
//     C[rows, N] (fp32) = sum_terms A_t[rows, K] . W_t[N, K]^T (+bias) (ReLU)  [+ per-group sum / sum^2]
// A operands are fp16 K-major and arrive through rank-3 TMA tensor maps {K, inner, outer}; the
// (inner, outer) box is 128 rows, which lets a conv1d(k=3) read its im2col rows straight out of a
// channel-last zero-padded activation buffer as OVERLAPPING windows (row stride = C elements) --
// no im2col pass.  With n_terms = 3 the product is evaluated as hi.hi + lo.hi + hi.lo of an
// (fp16 hi, fp16 lo) split of both operands: fp32-equivalent accuracy (2^-22) at 3 MMAs, fp32
// accumulation in TMEM.  Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue
// (TMEM -> registers -> global), 4-stage smem ring, two TMEM accumulators.
#include "tc_gemm.h"
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <vector>

namespace mind {
namespace tcg {

constexpr int kThreads = 320;      // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)
// operand ring: stages of (16 KB A + n_tile x 128 B W) -- 3 stages at n_tile = 256, 4 at 128, up to 8 below; in resident-W mode
// the W k-blocks sit at the front of the region and the rest is a ring of 16 KB A stages
constexpr uint32_t RING_BYTES = 147456;
constexpr uint32_t STAGE_A = 16384;
constexpr uint32_t RES_W_MAX = 98304;          // resident-W mode: W region <= 96 KB, leaving >= 3 A stages
// epilogue staging: each epilogue warp transposes its [32 rows x 32 columns] fp32 chunk through shared memory so that global
// stores leave as full 128-byte row segments (one row per thread straight out of TMEM is 32 sectors per store instruction)
constexpr uint32_t STG_STRIDE = 36;            // floats per staged row: 16-byte stores of 32 lanes and the transposed reads are conflict-free
constexpr uint32_t STG_WARP = 32 * STG_STRIDE * 4;
constexpr uint32_t SM_STG = RING_BYTES;
constexpr uint32_t SM_BAR = SM_STG + 8 * STG_WARP;
constexpr uint32_t SM_TMEM = SM_BAR + 192;
constexpr uint32_t SM_GN = SM_TMEM + 16;             // GroupNorm epilogue: gamma | beta | gamma2 | beta2 (4 x 256 floats), partial sums float2 [2 acc][8 warps][4 slots]
constexpr uint32_t SMEM_BYTES = SM_GN + 4096 + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            if (err) atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#define TCG_LD_X32(taddr, r)                                                                                       \
    asm volatile(                                                                                                  \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                  \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"  \
        "%28,%29,%30,%31}, [%32];"                                                                                 \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),         \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),   \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])  \
        : "r"(taddr) : "memory")

struct Args {
    int n_terms, k_blocks;
    int a_sel[3], w_k_off[3];
    int tiles_inner, tiles_outer, tiles_n;
    int r_in, r_out;
    int L_inner, n_outer;
    int N, n_tile;
    float* C; int ldc; int c_last;
    __half* Chi; __half* Clo; int ldh;      // optional fp16 (hi, lo) copy of the output (operand of a following GEMM)
    const float* bias; int relu;
    const float* gbias; int gsize, ldg;     // per row-group bias [(row / gsize), N]
    float* stats;
    int* err;
    int w_resident;            // all W k-blocks fit in smem: loaded once per CTA, A ring of 16 KB stages
    int n_stages; uint32_t stage_bytes;      // streaming mode ring
    int res_stages; uint32_t res_w_bytes;    // resident mode: A stages behind the W region
    int r_in_shift;                          // r_in is a power of two: tile row -> (outer, inner) by shift / mask
    // GroupNorm(1 group) epilogue (ActorNet): a tile holds whole outer groups (tiles_inner = 1), so the statistics over a
    // group's L_inner x N outputs are complete inside the CTA: y = GN(acc) (+ res) (ReLU) leaves as the padded fp16 (hi, lo)
    // operand of the next conv, no fp32 round trip through HBM and no second kernel
    int gn;
    const float* gn_gamma; const float* gn_beta; int gn_C;      // channel of column n: n & (gn_C - 1)
    float gn_inv_n;                                              // 1 / (outputs per group)
    const __half* gn_res_hi; const __half* gn_res_lo;            // identity shortcut (same padded layout) or null
    __half* gn_out_hi; __half* gn_out_lo;
    int64_t gn_ld_group;                                         // halfs per group in the padded layout: (L + 2) * C
    int gn_pad_tail;                                             // halfs of the last padded row: C
    // second GroupNorm'd input (the conv shortcut of a block's first Res1d): raw fp32 [group][L_inner][N] of another GEMM and
    // its statistics in that GEMM's `stats` layout (3 inner tiles x 2 column halves)
    const float* gn_res_raw; const float* gn_res_stats; const float* gn_res_gamma; const float* gn_res_beta;
    const float* gn_up_prev;                                     // FPN top-down step: + linear x2 upsampling of [group][L/2][C] fp32
    float* gn_out_f32;                                           // fp32 output [group][L_inner][N] instead of the (hi, lo) operand
};

// Epilogue of the GroupNorm mode (8 epilogue warps; warp w: TMEM lane quadrant lg = w & 3, column half (w - 2) >> 2).
// Pass 1 sums x and x^2 of the thread's row over its column half, reduces over the rows of the group inside the warp and
// publishes one partial per (warp, group slot); one 256-thread named barrier later every thread adds the partials of its
// group in a fixed order.  Pass 2 re-reads the accumulator (TMEM reads are cheap), normalises, adds the shortcut, applies
// ReLU, splits to fp16 (hi, lo) and leaves through the warp's staging tile as 64-byte row segments.
__device__ __forceinline__ void gn_epilogue(const Args& g, uint8_t* sgen, uint32_t bars, uint32_t tmem, int total_tiles, int warp, int lane) {
    float* sG = reinterpret_cast<float*>(sgen + SM_GN);
    float* sB = sG + 256;
    float* sG2 = sG + 512;
    float* sB2 = sG + 768;
    float2* sPart = reinterpret_cast<float2*>(sgen + SM_GN + 4096);
    for (int i = (int)threadIdx.x - 64; i < g.gn_C; i += 256) {
        sG[i] = g.gn_gamma[i]; sB[i] = g.gn_beta[i];
        if (g.gn_res_raw) { sG2[i] = g.gn_res_gamma[i]; sB2[i] = g.gn_res_beta[i]; }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    int acc = 0; uint32_t acc_phase = 0;
    const int ew = warp - 2, lg = warp & 3, chalf = ew >> 2;
    const int col_lo = chalf * (g.n_tile >> 1), col_hi = col_lo + (g.n_tile >> 1);
    const int row = lg * 32 + lane;
    const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
    const int rw = g.r_in < 32 ? g.r_in : 32;               // rows of one group inside a warp
    const int slot = g.r_in < 32 ? (lane >> g.r_in_shift) : 0;
    uint8_t* stg = sgen + SM_STG + (uint32_t)ew * STG_WARP; // [hi: 32 rows x 64 B][lo: 32 rows x 64 B], 16-byte chunks XOR-swizzled
    const int cmask = g.gn_C - 1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int outer = tile * g.r_out + (row >> g.r_in_shift), inner = row & (g.r_in - 1);
        const bool valid = outer < g.n_outer && inner < g.L_inner;
        mbar_wait(bars + 128 + 8 * acc, acc_phase, g.err, 14);
        tc_fence_after();
        float s1 = 0.f, s2 = 0.f;
        for (int n0 = col_lo; n0 < col_hi; n0 += 32) {
            uint32_t r[32];
            TCG_LD_X32(tmem + lane_base + acc * 256 + n0, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 32; ++e) { const float x = __uint_as_float(r[e]); s1 += x; s2 = fmaf(x, x, s2); }
        }
        if (!valid) { s1 = 0.f; s2 = 0.f; }
        for (int off = rw >> 1; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if ((lane & (rw - 1)) == 0) sPart[(acc * 8 + chalf * 4 + lg) * 4 + slot] = make_float2(s1, s2);   // (column half, lane quadrant): warp 2 is quadrant 2
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float S1, S2;
        if (g.r_in <= 32) {
            const float2 p0 = sPart[(acc * 8 + lg) * 4 + slot], p1 = sPart[(acc * 8 + 4 + lg) * 4 + slot];
            S1 = p0.x + p1.x; S2 = p0.y + p1.y;
        } else {        // r_in = 64: a group spans the lane quadrants 2a and 2a + 1
            const int l0 = lg & ~1;
            const float2 p0 = sPart[(acc * 8 + l0) * 4], p1 = sPart[(acc * 8 + l0 + 1) * 4];
            const float2 p2 = sPart[(acc * 8 + 4 + l0) * 4], p3 = sPart[(acc * 8 + 4 + l0 + 1) * 4];
            S1 = (p0.x + p1.x) + (p2.x + p3.x); S2 = (p0.y + p1.y) + (p2.y + p3.y);
        }
        const float mean = S1 * g.gn_inv_n;
        const float rstd = rsqrtf(fmaxf(S2 * g.gn_inv_n - mean * mean, 0.f) + 1e-5f);
        // this row in the padded layout: group base + one pad row + inner * N (N = fold * C consecutive halfs)
        const int64_t rbase = (int64_t)outer * g.gn_ld_group + g.gn_pad_tail + (int64_t)inner * g.N;
        const int64_t frow = ((int64_t)outer * g.L_inner + inner) * g.N;          // the same row in an unpadded fp32 [group][L_inner][N] tensor
        float m2 = 0.f, rstd2 = 0.f;
        if (g.gn_res_raw && valid) {          // statistics of the second input: partial sums in the other GEMM's layout, fixed order
            const float* q = g.gn_res_stats + (int64_t)outer * 12;
            m2 = (((q[0] + q[2]) + (q[4] + q[6])) + (q[8] + q[10])) * g.gn_inv_n;
            rstd2 = rsqrtf(fmaxf((((q[1] + q[3]) + (q[5] + q[7])) + (q[9] + q[11])) * g.gn_inv_n - m2 * m2, 0.f) + 1e-5f);
        }
        const float* up0 = nullptr; const float* up1 = nullptr; float lam = 0.f;
        if (g.gn_up_prev && valid) {          // interpolate(scale 2, linear, align_corners = False) source rows of this step
            const int Lp = g.L_inner >> 1;
            const float src = fmaxf(((float)inner + 0.5f) * 0.5f - 0.5f, 0.f);
            const int i0 = (int)floorf(src), i1 = min(i0 + 1, Lp - 1);
            lam = src - (float)i0;
            up0 = g.gn_up_prev + ((int64_t)outer * Lp + i0) * g.N;
            up1 = g.gn_up_prev + ((int64_t)outer * Lp + i1) * g.N;
        }
        for (int n0 = col_lo; n0 < col_hi; n0 += 32) {
            uint32_t r[32];
            TCG_LD_X32(tmem + lane_base + acc * 256 + n0, r);
            uint4 rh[4] = {}, rl[4] = {};
            if (g.gn_res_hi && valid) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    rh[k] = *reinterpret_cast<const uint4*>(g.gn_res_hi + rbase + n0 + k * 8);
                    rl[k] = *reinterpret_cast<const uint4*>(g.gn_res_lo + rbase + n0 + k * 8);
                }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // 8 columns -> one 16-byte chunk of hi and of lo (or two of fp32)
                float y[8];
                const uint32_t rhw[4] = {rh[k].x, rh[k].y, rh[k].z, rh[k].w}, rlw[4] = {rl[k].x, rl[k].y, rl[k].z, rl[k].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = (n0 + k * 8 + e * 2) & cmask;
                    const float2 gm = *reinterpret_cast<const float2*>(sG + c), bt = *reinterpret_cast<const float2*>(sB + c);
                    y[e * 2] = (__uint_as_float(r[k * 8 + e * 2]) - mean) * rstd * gm.x + bt.x;
                    y[e * 2 + 1] = (__uint_as_float(r[k * 8 + e * 2 + 1]) - mean) * rstd * gm.y + bt.y;
                    if (g.gn_res_hi && valid) {
                        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&rhw[e]));
                        const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&rlw[e]));
                        y[e * 2] += h.x + l.x; y[e * 2 + 1] += h.y + l.y;
                    }
                }
                if (g.gn_res_raw && valid) {
                    const float4 a0 = *reinterpret_cast<const float4*>(g.gn_res_raw + frow + n0 + k * 8);
                    const float4 a1 = *reinterpret_cast<const float4*>(g.gn_res_raw + frow + n0 + k * 8 + 4);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = (n0 + k * 8 + e) & cmask;
                        y[e] += (av[e] - m2) * rstd2 * sG2[c] + sB2[c];
                    }
                }
                if (up0) {
                    const float4 p0 = *reinterpret_cast<const float4*>(up0 + n0 + k * 8), p1 = *reinterpret_cast<const float4*>(up1 + n0 + k * 8);
                    const float4 p2 = *reinterpret_cast<const float4*>(up0 + n0 + k * 8 + 4), p3 = *reinterpret_cast<const float4*>(up1 + n0 + k * 8 + 4);
                    const float u0[8] = {p0.x, p0.y, p0.z, p0.w, p2.x, p2.y, p2.z, p2.w}, u1[8] = {p1.x, p1.y, p1.z, p1.w, p3.x, p3.y, p3.z, p3.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = u0[e] * (1.f - lam) + u1[e] * lam + y[e];
                }
                if (g.relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
                }
                if (g.gn_out_f32) {
                    *reinterpret_cast<float4*>(stg + (uint32_t)lane * 128u + (uint32_t)(((2 * k) ^ (lane & 7)) * 16)) = make_float4(y[0], y[1], y[2], y[3]);
                    *reinterpret_cast<float4*>(stg + (uint32_t)lane * 128u + (uint32_t)(((2 * k + 1) ^ (lane & 7)) * 16)) = make_float4(y[4], y[5], y[6], y[7]);
                } else {
                    uint32_t oh[4], ol[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 hh = __floats2half2_rn(y[e * 2], y[e * 2 + 1]);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(y[e * 2] - hf.x, y[e * 2 + 1] - hf.y);
                        oh[e] = *reinterpret_cast<const uint32_t*>(&hh); ol[e] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    const uint32_t so = (uint32_t)lane * 64u + (uint32_t)((k ^ ((lane >> 1) & 3)) * 16);
                    *reinterpret_cast<uint4*>(stg + so) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                    *reinterpret_cast<uint4*>(stg + 2048 + so) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                }
            }
            __syncwarp();
            if (g.gn_out_f32) {
                // write-out: 8 lanes cover the 128 bytes of one row's chunk, 4 rows per instruction
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3), ch = lane & 7;
                    const int trow = lg * 32 + rr;
                    const int outer_r = tile * g.r_out + (trow >> g.r_in_shift), inner_r = trow & (g.r_in - 1);
                    if (outer_r >= g.n_outer || inner_r >= g.L_inner) continue;
                    const float4 v = *reinterpret_cast<const float4*>(stg + (uint32_t)rr * 128u + (uint32_t)((ch ^ (rr & 7)) * 16));
                    *reinterpret_cast<float4*>(g.gn_out_f32 + ((int64_t)outer_r * g.L_inner + inner_r) * g.N + n0 + ch * 4) = v;
                }
            } else {
                // write-out: 4 lanes cover the 64 bytes of one row's chunk, 8 rows per instruction
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int rr = it * 8 + (lane >> 2), k = lane & 3;
                    const int trow = lg * 32 + rr;
                    const int outer_r = tile * g.r_out + (trow >> g.r_in_shift), inner_r = trow & (g.r_in - 1);
                    if (outer_r >= g.n_outer || inner_r >= g.L_inner) continue;
                    const uint32_t so = (uint32_t)rr * 64u + (uint32_t)((k ^ ((rr >> 1) & 3)) * 16);
                    const int64_t o = (int64_t)outer_r * g.gn_ld_group + g.gn_pad_tail + (int64_t)inner_r * g.N + n0 + k * 8;
                    *reinterpret_cast<uint4*>(g.gn_out_hi + o) = *reinterpret_cast<const uint4*>(stg + so);
                    *reinterpret_cast<uint4*>(g.gn_out_lo + o) = *reinterpret_cast<const uint4*>(stg + 2048 + so);
                }
            }
            __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 144 + 8 * acc);
        // zero pad rows (first and last row of every group this warp owns; hi by the first column half's warps, lo by the
        // second's) and, behind the last group, the rows a K-padded window of the next conv may read
        if (!g.gn_out_f32) {
            __half* dst = chalf ? g.gn_out_lo : g.gn_out_hi;
            const int n_slots = g.r_in < 32 ? (32 >> g.r_in_shift) : 1;
            if (g.r_in <= 32 || !(lg & 1)) {
                for (int sl = 0; sl < n_slots; ++sl) {
                    const int og = tile * g.r_out + ((lg * 32) >> g.r_in_shift) + sl;
                    if (og >= g.n_outer) break;
                    uint32_t* z0 = reinterpret_cast<uint32_t*>(dst + (int64_t)og * g.gn_ld_group);
                    uint32_t* z1 = reinterpret_cast<uint32_t*>(dst + (int64_t)(og + 1) * g.gn_ld_group - g.gn_pad_tail);
                    for (int c2 = lane; c2 < (g.gn_pad_tail >> 1); c2 += 32) { z0[c2] = 0u; z1[c2] = 0u; }
                    if (og == g.n_outer - 1) {
                        uint32_t* zt = reinterpret_cast<uint32_t*>(dst + (int64_t)g.n_outer * g.gn_ld_group);
                        for (int c2 = lane; c2 < 2 * g.gn_pad_tail; c2 += 32) zt[c2] = 0u;
                    }
                }
            }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap amap0, const __grid_constant__ CUtensorMap amap1,
          const __grid_constant__ CUtensorMap wmap, Args g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    volatile uint32_t* sTmem = reinterpret_cast<volatile uint32_t*>(sgen + SM_TMEM);
    const uint32_t bars = sbase + SM_BAR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        // full[s] = bars + 8 s, empty[s] = bars + 64 + 8 s (s < 8), tmem_full[a] = bars + 128 + 8 a,
        // tmem_empty[a] = bars + 144 + 8 a, W-resident = bars + 160
        for (int s = 0; s < 8; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 64 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bars + 128 + 8 * a, 1); mbar_init(bars + 144 + 8 * a, g.n_tile >= 64 ? 8 : 4); }
        mbar_init(bars + 160, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_TMEM), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const int total_tiles = g.tiles_outer * g.tiles_inner * g.tiles_n;
    const int nq = g.n_terms * g.k_blocks;

    // resident mode: smem = [W: all k-blocks, <= RES_W_MAX][A ring: res_stages x 16 KB]; bar_wres = bars + 160
    const uint32_t bar_wres = bars + 160;
    const int nkb_w = g.n_terms == 3 ? 2 * g.k_blocks : g.k_blocks;          // W k-blocks (hi | lo)
    const uint32_t wblk = (uint32_t)g.n_tile * 128u;                          // bytes of one W k-block
    const uint32_t res_a0 = sbase + g.res_w_bytes;                            // A ring base in resident mode
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            if (g.w_resident) {
                mbar_expect_tx(bar_wres, (uint32_t)nkb_w * wblk);
                for (int kb = 0; kb < nkb_w; ++kb) tma_load_2d(sbase + kb * wblk, &wmap, bar_wres, kb * 64, 0);
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int c = tile % g.tiles_inner, o = tile / g.tiles_inner;
                    for (int kb = 0; kb < g.k_blocks; ++kb)
                        for (int part = 0; part < (g.n_terms == 3 ? 2 : 1); ++part) {     // A_hi[kb], then A_lo[kb]
                            mbar_wait(bars + 64 + 8 * stage, phase ^ 1, g.err, 11);
                            const uint32_t full = bars + 8 * stage;
                            mbar_expect_tx(full, STAGE_A);
                            tma_load_3d(res_a0 + stage * STAGE_A, part ? &amap1 : &amap0, full, kb * 64, c * g.r_in, o * g.r_out);
                            if (++stage == g.res_stages) { stage = 0; phase ^= 1; }
                        }
                }
            } else {
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int nt = tile % g.tiles_n; const int rest = tile / g.tiles_n;
                    const int c = rest % g.tiles_inner, o = rest / g.tiles_inner;
                    for (int q = 0; q < nq; ++q) {
                        const int term = q / g.k_blocks, kb = q - term * g.k_blocks;
                        mbar_wait(bars + 64 + 8 * stage, phase ^ 1, g.err, 11);
                        const uint32_t full = bars + 8 * stage;
                        mbar_expect_tx(full, STAGE_A + (uint32_t)g.n_tile * 128u);
                        const uint32_t sA = sbase + stage * g.stage_bytes, sW = sA + STAGE_A;
                        tma_load_3d(sA, g.a_sel[term] ? &amap1 : &amap0, full, kb * 64, c * g.r_in, o * g.r_out);
                        tma_load_2d(sW, &wmap, full, g.w_k_off[term] + kb * 64, nt * g.n_tile);
                        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            const uint32_t idesc = umma_idesc_f16(g.n_tile);
            if (g.w_resident) {
                mbar_wait(bar_wres, 0, g.err, 15);
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    mbar_wait(bars + 144 + 8 * acc, acc_phase ^ 1, g.err, 12);
                    tc_fence_after();
                    uint32_t first = 1;
                    for (int kb = 0; kb < g.k_blocks; ++kb)
                        for (int part = 0; part < (g.n_terms == 3 ? 2 : 1); ++part) {
                            mbar_wait(bars + 8 * stage, phase, g.err, 13);
                            tc_fence_after();
                            const uint32_t sA = res_a0 + stage * STAGE_A;
                            const uint32_t sWhi = sbase + kb * wblk, sWlo = sbase + (g.k_blocks + kb) * wblk;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {       // A_hi: x W_hi and x W_lo ; A_lo: x W_hi
                                umma_f16(tmem + acc * 256, umma_desc_sw128(sA + kk * 32), umma_desc_sw128(sWhi + kk * 32), idesc, first ? 0u : 1u);
                                first = 0;
                            }
                            if (part == 0 && g.n_terms == 3) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_f16(tmem + acc * 256, umma_desc_sw128(sA + kk * 32), umma_desc_sw128(sWlo + kk * 32), idesc, 1u);
                            }
                            umma_commit(bars + 64 + 8 * stage);
                            if (++stage == g.res_stages) { stage = 0; phase ^= 1; }
                        }
                    umma_commit(bars + 128 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            } else {
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    mbar_wait(bars + 144 + 8 * acc, acc_phase ^ 1, g.err, 12);
                    tc_fence_after();
                    for (int q = 0; q < nq; ++q) {
                        mbar_wait(bars + 8 * stage, phase, g.err, 13);
                        tc_fence_after();
                        const uint32_t sA = sbase + stage * g.stage_bytes, sW = sA + STAGE_A;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_f16(tmem + acc * 256, umma_desc_sw128(sA + kk * 32), umma_desc_sw128(sW + kk * 32), idesc,
                                     (q | kk) != 0);
                        umma_commit(bars + 64 + 8 * stage);
                        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(bars + 128 + 8 * acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (g.gn) {
        gn_epilogue(g, sgen, bars, tmem, total_tiles, warp, lane);
    } else {
        int acc = 0; uint32_t acc_phase = 0;
        const int lg = warp & 3;                       // TMEM lane quadrant (warp id mod 4)
        const int chalf = (warp - 2) >> 2;             // warps 2-5: first column half, 6-9: second half
        const bool split_cols = g.n_tile >= 64;
        const int col_lo = split_cols ? chalf * (g.n_tile >> 1) : 0;
        const int col_hi = split_cols ? col_lo + (g.n_tile >> 1) : g.n_tile;
        const int row = lg * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
        for (int tile = blockIdx.x; tile < total_tiles && (split_cols || chalf == 0); tile += gridDim.x) {
            const int nt = tile % g.tiles_n; const int rest = tile / g.tiles_n;
            const int c = rest % g.tiles_inner, o = rest / g.tiles_inner;
            const int outer = o * g.r_out + (row >> g.r_in_shift), inner = c * g.r_in + (row & (g.r_in - 1));
            const bool valid = outer < g.n_outer && inner < g.L_inner;
            const int64_t orow = (int64_t)outer * g.L_inner + inner;
            const float* grow = (g.gbias && valid) ? g.gbias + (orow / g.gsize) * g.ldg : nullptr;
            mbar_wait(bars + 128 + 8 * acc, acc_phase, g.err, 14);
            tc_fence_after();
            float s1 = 0.f, s2 = 0.f;
            for (int n0 = col_lo; n0 < col_hi; n0 += 32) {
                uint32_t r[32];
                TCG_LD_X32(tmem + lane_base + acc * 256 + n0, r);
                const int nb = nt * g.n_tile + n0;
                // bias / group-bias for the whole 32-column chunk first (vector loads, independent of the TMEM data)
                float4 bz[8];
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const int n = nb + k4 * 4;
                    bz[k4] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n + 3 < g.N) {
                        if (g.bias) bz[k4] = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                        if (grow) {
                            const float4 gz = __ldg(reinterpret_cast<const float4*>(grow + n));
                            bz[k4].x += gz.x; bz[k4].y += gz.y; bz[k4].z += gz.z; bz[k4].w += gz.w;
                        }
                    } else {
                        float t4[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int e = 0; e < 4; ++e)
                            if (n + e < g.N) t4[e] = (g.bias ? __ldg(g.bias + n + e) : 0.f) + (grow ? __ldg(grow + n + e) : 0.f);
                        bz[k4] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* stg = reinterpret_cast<float*>(sgen + SM_STG + (uint32_t)(warp - 2) * STG_WARP);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float bb[4] = {bz[k4].x, bz[k4].y, bz[k4].z, bz[k4].w};
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int n = nb + k4 * 4 + e;
                        float x = __uint_as_float(r[k4 * 4 + e]) + bb[e];
                        if (g.relu) x = fmaxf(x, 0.f);
                        if (n < g.N) { s1 += x; s2 += x * x; }
                        v[e] = x;
                    }
                    *reinterpret_cast<float4*>(stg + lane * STG_STRIDE + k4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
                }
                __syncwarp();
                // transposed write-out: 8 lanes cover the 32 columns of one row, 4 rows per instruction
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3), c4 = lane & 7;
                    const int trow = lg * 32 + rr;
                    const int outer_r = o * g.r_out + (trow >> g.r_in_shift), inner_r = c * g.r_in + (trow & (g.r_in - 1));
                    if (outer_r >= g.n_outer || inner_r >= g.L_inner) continue;
                    if (g.c_last && inner_r != g.L_inner - 1) continue;          // only the last inner row is kept (statistics over all)
                    const float4 x = *reinterpret_cast<const float4*>(stg + rr * STG_STRIDE + c4 * 4);
                    const int64_t orow_r = g.c_last ? (int64_t)outer_r : (int64_t)outer_r * g.L_inner + inner_r;
                    const int n = nb + c4 * 4;
                    if (g.C) {
                        float* cr = g.C + orow_r * g.ldc;
                        if (n + 3 < g.N && ((g.ldc & 3) == 0)) {
                            *reinterpret_cast<float4*>(cr + n) = x;
                        } else {
                            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (n + e < g.N) cr[n + e] = xv[e];
                        }
                    }
                    if (g.Chi && n + 3 < g.N) {
                        const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
                        uint2 uh, ul;
                        uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
                        ul.x = *reinterpret_cast<const uint32_t*>(&l01); ul.y = *reinterpret_cast<const uint32_t*>(&l23);
                        *reinterpret_cast<uint2*>(g.Chi + orow_r * g.ldh + n) = uh;
                        *reinterpret_cast<uint2*>(g.Clo + orow_r * g.ldh + n) = ul;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 144 + 8 * acc);
            if (g.stats) {   // per-(outer, inner tile) partial sums; r_in is a power of two <= 32
                if (!valid) { s1 = 0.f; s2 = 0.f; }
                for (int off = g.r_in >> 1; off > 0; off >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                }
                if ((row % g.r_in) == 0 && outer < g.n_outer) {   // [outer][inner tile][column half][2]
                    float* sp = g.stats + (((int64_t)outer * g.tiles_inner + c) * 2 + chalf) * 2;
                    sp[0] = s1; sp[1] = s2;
                    if (!split_cols) { sp[2] = 0.f; sp[3] = 0.f; }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------
// element-wise companions of the conv GEMMs (channel-last, zero-padded fp16 hi/lo activations)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store(__half* hi, __half* lo, int64_t idx, float x) {
    const __half h = __float2half_rn(x);
    hi[idx] = h;
    lo[idx] = __float2half_rn(x - __half2float(h));
}

// actors [A,14,48] fp32 (channel-first) -> hi/lo [A][50][16] channel-last with zero pad rows / channels
__global__ void k_actor_prep(const float* __restrict__ actors, __half* __restrict__ hi, __half* __restrict__ lo, int A) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)A * 50 * 16) return;
    const int c = (int)(idx & 15);
    const int t = (int)((idx >> 4) % 50);
    const int a = (int)(idx / 800);
    float x = 0.f;
    if (c < 14 && t >= 1 && t <= 48) x = actors[((int64_t)a * 14 + c) * 48 + (t - 1)];
    split_store(hi, lo, idx, x);
}

// y = GN(raw) (+ GN(res_raw) | + (res_hi+res_lo))  (ReLU)  -> hi/lo padded [A][L+2][C] and / or fp32 [A][L][C]
// stats: [A][3 inner tiles][2 column halves][2] partial (sum, sum^2) over the actor's C*L outputs
struct ApplyArgs {
    const float* raw; const float* stats; const float* gamma; const float* beta;
    const float* res_raw; const float* res_stats; const float* res_gamma; const float* res_beta;
    const __half* res_hi; const __half* res_lo;     // identity shortcut, padded [A][L+2][C]
    __half* out_hi; __half* out_lo; float* out_f32;
    const float* up_prev;   // + linear x2 upsampling of [A][L/2][C] (FPN top-down step, network.py:57-58)
    int last_only;          // raw / out_f32 hold time step L-1 only
    int A, L, C, relu;
    int c_shift;            // log2(C) when C is a power of two, else -1
};
__global__ void __launch_bounds__(256) k_gn_apply(ApplyArgs p) {
    // one WARP per actor: every lane forms the statistics from the 12 (+12) partial sums itself (broadcast loads), so there is
    // no single-thread phase and no block barrier in front of a few hundred elements of work
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int n = p.L * p.C, n4 = n >> 2;
    const float fn = (float)n;
    for (int a = blockIdx.x * wpb + (threadIdx.x >> 5); a < p.A; a += gridDim.x * wpb) {
        const float* s = p.stats + (int64_t)a * 12;     // 3 inner tiles x 2 column halves x (sum, sum^2), fixed order
        const float mean = (((s[0] + s[2]) + (s[4] + s[6])) + (s[8] + s[10])) / fn;
        const float var = fmaxf((((s[1] + s[3]) + (s[5] + s[7])) + (s[9] + s[11])) / fn - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        float m2 = 0.f, rstd2 = 0.f;
        if (p.res_raw) {
            const float* r = p.res_stats + (int64_t)a * 12;
            m2 = (((r[0] + r[2]) + (r[4] + r[6])) + (r[8] + r[10])) / fn;
            const float v2 = fmaxf((((r[1] + r[3]) + (r[5] + r[7])) + (r[9] + r[11])) / fn - m2 * m2, 0.f);
            rstd2 = rsqrtf(v2 + 1e-5f);
        }
        const int first4 = p.last_only ? ((p.L - 1) * p.C) >> 2 : 0;
        const int64_t rbase = p.last_only ? (int64_t)a * p.C - (int64_t)(p.L - 1) * p.C : (int64_t)a * n;   // raw / out_f32 row base
        for (int i4 = first4 + lane; i4 < n4; i4 += 32) {      // C % 4 == 0: 4 channels per iteration
            const int i = i4 << 2;
            const int t = p.c_shift >= 0 ? (i >> p.c_shift) : i / p.C, c = i - t * p.C;
            const float4 x = *reinterpret_cast<const float4*>(p.raw + rbase + i);
            const float4 gm = *reinterpret_cast<const float4*>(p.gamma + c), bt = *reinterpret_cast<const float4*>(p.beta + c);
            float y[4] = {(x.x - mean) * rstd * gm.x + bt.x, (x.y - mean) * rstd * gm.y + bt.y,
                          (x.z - mean) * rstd * gm.z + bt.z, (x.w - mean) * rstd * gm.w + bt.w};
            const int64_t pidx = ((int64_t)a * (p.L + 2) + t + 1) * p.C + c;
            if (p.res_raw) {
                const float4 r = *reinterpret_cast<const float4*>(p.res_raw + (int64_t)a * n + i);
                const float4 g2 = *reinterpret_cast<const float4*>(p.res_gamma + c), b2 = *reinterpret_cast<const float4*>(p.res_beta + c);
                y[0] += (r.x - m2) * rstd2 * g2.x + b2.x; y[1] += (r.y - m2) * rstd2 * g2.y + b2.y;
                y[2] += (r.z - m2) * rstd2 * g2.z + b2.z; y[3] += (r.w - m2) * rstd2 * g2.w + b2.w;
            } else if (p.res_hi) {
                const uint2 h = *reinterpret_cast<const uint2*>(p.res_hi + pidx), l = *reinterpret_cast<const uint2*>(p.res_lo + pidx);
                const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
                const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
                y[0] += h0.x + l0.x; y[1] += h0.y + l0.y; y[2] += h1.x + l1.x; y[3] += h1.y + l1.y;
            }
            if (p.up_prev) {      // out = lerp_x2(prev) + lateral, same expression as the reference's interpolate(align_corners=False)
                const int Lp = p.L >> 1;
                const float src = fmaxf(((float)t + 0.5f) * 0.5f - 0.5f, 0.f);
                const int i0 = (int)floorf(src), i1 = min(i0 + 1, Lp - 1);
                const float lam = src - (float)i0;
                const float* pp = p.up_prev + (int64_t)a * Lp * p.C;
                const float4 p0 = *reinterpret_cast<const float4*>(pp + i0 * p.C + c), p1 = *reinterpret_cast<const float4*>(pp + i1 * p.C + c);
                y[0] = p0.x * (1.f - lam) + p1.x * lam + y[0]; y[1] = p0.y * (1.f - lam) + p1.y * lam + y[1];
                y[2] = p0.z * (1.f - lam) + p1.z * lam + y[2]; y[3] = p0.w * (1.f - lam) + p1.w * lam + y[3];
            }
            if (p.relu) { y[0] = fmaxf(y[0], 0.f); y[1] = fmaxf(y[1], 0.f); y[2] = fmaxf(y[2], 0.f); y[3] = fmaxf(y[3], 0.f); }
            if (p.out_hi) {
                const __half2 a01 = __floats2half2_rn(y[0], y[1]), a23 = __floats2half2_rn(y[2], y[3]);
                const float2 f01 = __half22float2(a01), f23 = __half22float2(a23);
                const __half2 b01 = __floats2half2_rn(y[0] - f01.x, y[1] - f01.y), b23 = __floats2half2_rn(y[2] - f23.x, y[3] - f23.y);
                uint2 uh, ul;
                uh.x = *reinterpret_cast<const uint32_t*>(&a01); uh.y = *reinterpret_cast<const uint32_t*>(&a23);
                ul.x = *reinterpret_cast<const uint32_t*>(&b01); ul.y = *reinterpret_cast<const uint32_t*>(&b23);
                *reinterpret_cast<uint2*>(p.out_hi + pidx) = uh;
                *reinterpret_cast<uint2*>(p.out_lo + pidx) = ul;
            }
            if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + rbase + i) = make_float4(y[0], y[1], y[2], y[3]);
        }
        if (p.out_hi) {   // zero pad rows t = 0 and t = L+1
            for (int c = lane; c < 2 * p.C; c += 32) {
                const int64_t pidx = ((int64_t)a * (p.L + 2) + (c < p.C ? 0 : p.L + 1)) * p.C + (c < p.C ? c : c - p.C);
                p.out_hi[pidx] = __float2half(0.f);
                p.out_lo[pidx] = __float2half(0.f);
            }
            // K-padded conv windows of the LAST actor read up to 4 rows past the logical end of this
            // (possibly re-used, larger) buffer: keep them finite (they meet zero weights)
            if (a == p.A - 1) {
                const int64_t end = (int64_t)p.A * (p.L + 2) * p.C;
                for (int c = lane; c < 4 * p.C; c += 32) {
                    p.out_hi[end + c] = __float2half(0.f);
                    p.out_lo[end + c] = __float2half(0.f);
                }
            }
        }
    }
}

}  // namespace tcg

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
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
char g_err[256];
}  // namespace

const char* tcg_encode_a(void* map, const __half* base, int64_t k_extent, int64_t inner, int64_t outer,
                         int64_t inner_stride_elems, int64_t outer_stride_elems, int r_in, int r_out) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled unavailable";
    cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[2] = {(cuuint64_t)inner_stride_elems * 2, (cuuint64_t)outer_stride_elems * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)r_in, (cuuint32_t)r_out};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "encode A map failed (%d): k=%lld in=%lld out=%lld s=%lld/%lld box=%d/%d", (int)r,
                                      (long long)k_extent, (long long)inner, (long long)outer, (long long)inner_stride_elems,
                                      (long long)outer_stride_elems, r_in, r_out); return g_err; }
    return nullptr;
}

const char* tcg_encode_w(void* map, const __half* base, int64_t k_total, int64_t n_rows, int n_tile) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled unavailable";
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)n_tile};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "encode W map failed (%d)", (int)r); return g_err; }
    return nullptr;
}

const char* tcg_launch(const TcGemm& p, int sm_count, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tcg::k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg::SMEM_BYTES) != cudaSuccess)
            return "cudaFuncSetAttribute(tc_gemm) failed";
        attr = true;
    }
    if (p.r_in * p.r_out != 128 || (p.r_in & (p.r_in - 1))) return "tc_gemm: r_in * r_out must be 128, r_in a power of two";
    if (p.n_tile % 32 || p.n_tile > 256 || p.n_tile < 32) return "tc_gemm: n_tile must be a multiple of 32 in [32,256]";
    tcg::Args g;
    g.n_terms = p.split ? 3 : 1;
    g.k_blocks = p.k_blocks;
    g.a_sel[0] = 0; g.a_sel[1] = 1; g.a_sel[2] = 0;
    g.w_k_off[0] = 0; g.w_k_off[1] = 0; g.w_k_off[2] = p.k_blocks * 64;
    g.r_in = p.r_in; g.r_out = p.r_out; g.L_inner = p.L_inner; g.n_outer = p.n_outer;
    g.tiles_inner = (p.L_inner + p.r_in - 1) / p.r_in;
    g.tiles_outer = (p.n_outer + p.r_out - 1) / p.r_out;
    g.N = p.N; g.n_tile = p.n_tile; g.tiles_n = (p.N + p.n_tile - 1) / p.n_tile;
    g.C = p.C; g.ldc = p.ldc; g.c_last = p.c_last_only; g.Chi = p.Chi; g.Clo = p.Clo; g.ldh = p.ldh; g.bias = p.bias; g.relu = p.relu; g.stats = p.stats; g.err = p.err;
    g.gbias = p.gbias; g.gsize = p.gsize > 0 ? p.gsize : 1; g.ldg = p.ldg;
    g.gn = (p.gn_out_hi || p.gn_out_f32) ? 1 : 0;
    g.gn_res_raw = p.gn_res_raw; g.gn_res_stats = p.gn_res_stats; g.gn_res_gamma = p.gn_res_gamma; g.gn_res_beta = p.gn_res_beta;
    g.gn_up_prev = p.gn_up_prev; g.gn_out_f32 = p.gn_out_f32;
    g.gn_gamma = p.gn_gamma; g.gn_beta = p.gn_beta; g.gn_C = p.gn_C; g.gn_inv_n = p.gn_inv_n;
    g.gn_res_hi = p.gn_res_hi; g.gn_res_lo = p.gn_res_lo; g.gn_out_hi = p.gn_out_hi; g.gn_out_lo = p.gn_out_lo;
    g.gn_ld_group = p.gn_ld_group; g.gn_pad_tail = p.gn_C;
    if (g.gn) {
        if (g.tiles_inner != 1 || g.tiles_n != 1 || p.n_tile < 64 || p.n_tile != p.N || (p.gn_C & (p.gn_C - 1)) || p.gn_C > 256 || p.r_in > 64)
            return "tc_gemm: GroupNorm epilogue needs whole groups per tile, one column tile of 64..256 and a power-of-two channel count";
    }
    const int nkb_w = (p.split ? 2 : 1) * p.k_blocks;
    g.res_w_bytes = (uint32_t)nkb_w * (uint32_t)p.n_tile * 128u;
    g.w_resident = (g.tiles_n == 1 && (int64_t)nkb_w * p.n_tile * 128 <= (int64_t)tcg::RES_W_MAX) ? 1 : 0;
    g.res_stages = g.w_resident ? std::min<int>(8, (int)((tcg::RING_BYTES - g.res_w_bytes) / tcg::STAGE_A)) : 0;
    g.r_in_shift = 0;
    while ((1 << g.r_in_shift) < p.r_in) ++g.r_in_shift;
    g.stage_bytes = tcg::STAGE_A + (uint32_t)p.n_tile * 128u;
    g.n_stages = std::min<int>(8, (int)(tcg::RING_BYTES / g.stage_bytes));
    const int total = g.tiles_inner * g.tiles_outer * g.tiles_n;
    if (total <= 0) return nullptr;
    CUtensorMap a0, a1, w;
    memcpy(&a0, p.amap_hi, sizeof a0);
    memcpy(&a1, p.split ? p.amap_lo : p.amap_hi, sizeof a1);
    memcpy(&w, p.wmap, sizeof w);
    tcg::k_tc_gemm<<<std::min(total, sm_count), tcg::kThreads, tcg::SMEM_BYTES, st>>>(a0, a1, w, g);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

void tcg_actor_prep(const float* actors, __half* hi, __half* lo, int A, cudaStream_t st) {
    const int64_t n = (int64_t)A * 800;
    tcg::k_actor_prep<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(actors, hi, lo, A);
    ++g_launches;
}
void tcg_gn_apply(const TcApply& q, cudaStream_t st) {
    tcg::ApplyArgs p;
    p.raw = q.raw; p.stats = q.stats; p.gamma = q.gamma; p.beta = q.beta;
    p.res_raw = q.res_raw; p.res_stats = q.res_stats; p.res_gamma = q.res_gamma; p.res_beta = q.res_beta;
    p.res_hi = q.res_hi; p.res_lo = q.res_lo; p.out_hi = q.out_hi; p.out_lo = q.out_lo; p.out_f32 = q.out_f32;
    p.up_prev = q.up_prev; p.last_only = q.last_only;
    p.A = q.A; p.L = q.L; p.C = q.C; p.relu = q.relu;
    p.c_shift = -1;
    for (int sft = 0; sft < 12; ++sft) if ((1 << sft) == q.C) p.c_shift = sft;
    if (q.A <= 0) return;
    tcg::k_gn_apply<<<(unsigned)std::min((q.A + 7) / 8, 148 * 8), 256, 0, st>>>(p);
    ++g_launches;
}
}  // namespace mind
