// fp32 SIMT kernels of the MIND hot path (sm_100a).
//
// These are the exact-arithmetic kernels: encoders, node-side projections, decoder, and the
// un-fused fp32 version of the rela-fusion pair pipeline that serves as the on-device comparator
// for the tcgen05 kernel in fusion_tc.cu.  Reference semantics: planners/mind/networks/network.py
// and layers.py (line numbers cited per kernel).
#include "kernels.h"
#include <math.h>
#include <algorithm>

namespace mind {

int64_t g_launches = 0;

#define LN_EPS 1e-5f

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------
// generic C = A . W^T (+bias)(+group bias)(ReLU) ; 64x64x16 tiles, 256 threads, 4x4 per thread
// ------------------------------------------------------------------------------------------
constexpr int GBM = 64, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256) k_gemm_tn(GemmArgs g) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Ws[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * GBM;
    const int n0 = blockIdx.y * GBN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lrow = tid >> 2;          // 0..63
    const int lcol = (tid & 3) * 4;     // 0,4,8,12
    const bool a_vec = ((g.lda & 3) == 0) && ((((uintptr_t)g.A) & 15) == 0);
    const bool w_vec = ((g.ldw & 3) == 0) && ((((uintptr_t)g.W) & 15) == 0);

    const int kslice = g.ksplit > 1 ? ((g.K / g.ksplit + GBK - 1) / GBK) * GBK : g.K;
    const int k_begin = g.ksplit > 1 ? (int)blockIdx.z * kslice : 0, k_end = min(g.K, k_begin + kslice);
    for (int k0 = k_begin; k0 < k_end; k0 += GBK) {
        {   // A tile
            const int64_t m = m0 + lrow;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < g.M) {
                const float* p = g.A + m * (int64_t)g.lda + k0 + lcol;
                if (a_vec && k0 + lcol + 3 < k_end) {
                    float4 t = *reinterpret_cast<const float4*>(p);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (k0 + lcol + e < k_end) v[e] = p[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) As[lcol + e][lrow] = v[e];
        }
        {   // W tile
            const int n = n0 + lrow;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < g.N) {
                const float* p = g.W + (int64_t)n * g.ldw + k0 + lcol;
                if (w_vec && k0 + lcol + 3 < k_end) {
                    float4 t = *reinterpret_cast<const float4*>(p);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (k0 + lcol + e < k_end) v[e] = p[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) Ws[lcol + e][lrow] = v[e];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.ksplit > 1) { g.part[((int64_t)blockIdx.z * g.M + m) * g.N + n] = v; continue; }
            if (g.bias) v += g.bias[n];
            if (g.gbias) v += g.gbias[(m / g.gsize) * (int64_t)g.ldg + n];
            if (g.relu) v = fmaxf(v, 0.f);
            g.C[m * (int64_t)g.ldc + n] = v;
        }
    }
}

// second half of a split-K product: partial sums added in slice order, then bias / group bias / ReLU
__global__ void __launch_bounds__(256) k_splitk_reduce(GemmArgs g) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)g.M * g.N) return;
    const int64_t m = idx / g.N;
    const int n = (int)(idx - m * g.N);
    float v = 0.f;
    for (int z = 0; z < g.ksplit; ++z) v += g.part[(int64_t)z * g.M * g.N + idx];
    if (g.bias) v += g.bias[n];
    if (g.gbias) v += g.gbias[(m / g.gsize) * (int64_t)g.ldg + n];
    if (g.relu) v = fmaxf(v, 0.f);
    g.C[m * (int64_t)g.ldc + n] = v;
}

// large aligned case (K % 16 == 0, N % 128 == 0, 16-byte aligned rows): 128x128x16 tiles, 8x8 per
// thread, register-staged double buffering of the global loads
constexpr int HBM_ = 128, HBN_ = 128, HBK_ = 16;
__global__ void __launch_bounds__(256) k_gemm_tn_big(GemmArgs g) {
    __shared__ __align__(16) float As[2][HBK_][HBM_ + 4];
    __shared__ __align__(16) float Ws[2][HBK_][HBN_ + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;             // 16 x 16 threads, 8 x 8 outputs each
    const int64_t m0 = (int64_t)blockIdx.x * HBM_;
    const int n0 = blockIdx.y * HBN_;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    // each thread stages 2 float4 of A and 2 of W per k-step: rows lrow and lrow+64, k offset lcol
    const int lrow = tid >> 2, lcol = (tid & 3) * 4;
    float4 ra[2], rw[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t m = m0 + lrow + h * 64;
            ra[h] = (m < g.M) ? *reinterpret_cast<const float4*>(g.A + m * (int64_t)g.lda + k0 + lcol) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int n = n0 + lrow + h * 64;
            rw[h] = *reinterpret_cast<const float4*>(g.W + (int64_t)n * g.ldw + k0 + lcol);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + h * 64;
            As[buf][lcol + 0][r] = ra[h].x; As[buf][lcol + 1][r] = ra[h].y; As[buf][lcol + 2][r] = ra[h].z; As[buf][lcol + 3][r] = ra[h].w;
            Ws[buf][lcol + 0][r] = rw[h].x; Ws[buf][lcol + 1][r] = rw[h].y; Ws[buf][lcol + 2][r] = rw[h].z; Ws[buf][lcol + 3][r] = rw[h].w;
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    const int nk = g.K / HBK_;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * HBK_);
#pragma unroll
        for (int kk = 0; kk < HBK_; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x = acc[i][jh * 4 + e];
                if (g.bias) x += g.bias[n + e];
                if (g.gbias) x += g.gbias[(m / g.gsize) * (int64_t)g.ldg + n + e];
                if (g.relu) x = fmaxf(x, 0.f);
                v[e] = x;
            }
            *reinterpret_cast<float4*>(g.C + m * (int64_t)g.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// few rows (tree-level batches: M = 6 modes x a handful of actors): one warp per output column, lanes split K, rows in
// register chunks of 16.  The 64x64 tiled kernel above runs one or two CTAs through a serial K loop for these shapes
// (25 us per launch measured at M = 48); this one spreads the N columns over the chip.
__global__ void __launch_bounds__(128) k_gemm_small_m(GemmArgs g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 4 + warp;
    if (n >= g.N) return;
    const float* __restrict__ wrow = g.W + (int64_t)n * g.ldw;
    const int m_base = blockIdx.y * 64;
    const int m_end = min(g.M, m_base + 64);
    const float bn = g.bias ? g.bias[n] : 0.f;
    for (int m0 = m_base; m0 < m_end; m0 += 16) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int k = lane * 4; k < g.K; k += 128) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (m0 + i < m_end) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)(m0 + i) * g.lda + k));
                    acc[i] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[i]))));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float v0 = warp_sum(acc[i]);
            const int m = m0 + i;
            if (lane == 0 && m < m_end) {
                float v = v0 + bn;
                if (g.gbias) v += g.gbias[(m / g.gsize) * (int64_t)g.ldg + n];
                if (g.relu) v = fmaxf(v, 0.f);
                g.C[(int64_t)m * g.ldc + n] = v;
            }
        }
    }
}

void launch_gemm(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return;
    if (g.M <= 256 && (g.K & 3) == 0 && (g.lda & 3) == 0 && (g.ldw & 3) == 0 &&
        ((((uintptr_t)g.A) | ((uintptr_t)g.W)) & 15) == 0) {
        dim3 grid((unsigned)((g.N + 3) / 4), (unsigned)((g.M + 63) / 64));
        k_gemm_small_m<<<grid, 128, 0, st>>>(g);
        ++g_launches;
        return;
    }
    const bool big = g.M >= 2048 && (g.N % HBN_) == 0 && (g.K % HBK_) == 0 && (g.lda & 3) == 0 && (g.ldw & 3) == 0 &&
                     (g.ldc & 3) == 0 && ((((uintptr_t)g.A) | ((uintptr_t)g.W) | ((uintptr_t)g.C)) & 15) == 0;
    if (big) {
        dim3 grid((unsigned)((g.M + HBM_ - 1) / HBM_), (unsigned)(g.N / HBN_));
        k_gemm_tn_big<<<grid, 256, 0, st>>>(g);
    } else {
        dim3 grid((unsigned)((g.M + GBM - 1) / GBM), (unsigned)((g.N + GBN - 1) / GBN));
        // skinny output, long K (e.g. the decoder's 1536 -> 128 linear on B*6 rows: 48 CTAs walking K = 1536 serially):
        // split K over grid.z so that the chip is covered, reduce deterministically
        // scratch per device, grow-only and never freed: a captured CUDA graph may hold the pointer of an earlier, smaller one
        static float* s_part[16] = {nullptr};
        static size_t s_part_floats[16] = {0};
        const int ctas = (int)(grid.x * grid.y);
        int ks = 1, dev = 0;
        if (ctas < 96 && g.K >= 512) ks = std::min(std::min(8, g.K / 128), std::max(1, 296 / ctas));
        if (ks > 1 && (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16)) ks = 1;
        if (ks > 1) {
            const size_t need = (size_t)ks * (size_t)g.M * (size_t)g.N;
            if (need > s_part_floats[dev]) {     // first use of a shape is never inside a graph capture (mind_forward runs it plainly first)
                float* nb = nullptr;
                const size_t cap = std::max(need, 2 * s_part_floats[dev]);
                if (cudaMalloc(&nb, cap * sizeof(float)) == cudaSuccess) { s_part[dev] = nb; s_part_floats[dev] = cap; } else { cudaGetLastError(); ks = 1; }
            }
        }
        if (ks > 1) {
            GemmArgs h = g;
            h.ksplit = ks; h.part = s_part[dev];
            grid.z = (unsigned)ks;
            k_gemm_tn<<<grid, 256, 0, st>>>(h);
            k_splitk_reduce<<<(unsigned)(((int64_t)g.M * g.N + 255) / 256), 256, 0, st>>>(h);
            ++g_launches;
        } else {
            k_gemm_tn<<<grid, 256, 0, st>>>(g);
        }
    }
    ++g_launches;
}

// ------------------------------------------------------------------------------------------
// row LayerNorm (biased variance, eps 1e-5), optional residual input and ReLU; warp per row
// ------------------------------------------------------------------------------------------
template <int MAXV>   // MAXV float4 per lane: D <= MAXV*128
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ x, const float* __restrict__ res,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   float* __restrict__ out, int64_t rows, int D, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nv = D >> 2;   // float4 per row
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c4 = i * 32 + lane;
        if (c4 < nv) {
            float4 t = reinterpret_cast<const float4*>(x + row * D)[c4];
            if (res) {
                const float4 r = reinterpret_cast<const float4*>(res + row * D)[c4];
                t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
            }
            v[i] = t;
            s += (t.x + t.y) + (t.z + t.w);
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c4 = i * 32 + lane;
        if (c4 < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + LN_EPS);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c4 = i * 32 + lane;
        if (c4 < nv) {
            const float4 gm = reinterpret_cast<const float4*>(gamma)[c4];
            const float4 bt = reinterpret_cast<const float4*>(beta)[c4];
            float4 o;
            o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
            o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
            o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
            o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            reinterpret_cast<float4*>(out + row * D)[c4] = o;
        }
    }
}

void launch_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float* out,
                      int64_t rows, int D, int relu, cudaStream_t st) {
    if (rows <= 0) return;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (D <= 128) k_layernorm<1><<<grid, 256, 0, st>>>(x, res, gamma, beta, out, rows, D, relu);
    else if (D <= 768) k_layernorm<6><<<grid, 256, 0, st>>>(x, res, gamma, beta, out, rows, D, relu);
    else k_layernorm<12><<<grid, 256, 0, st>>>(x, res, gamma, beta, out, rows, D, relu);
    ++g_launches;
}

// LayerNorm over D = 128 that also emits the fp16 (hi, lo) split consumed by the tensor-core GEMM engine
__global__ void __launch_bounds__(256) k_layernorm_hl(const float* __restrict__ x, const float* __restrict__ res,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ out32, __half* __restrict__ hi, __half* __restrict__ lo,
                                                      int64_t rows, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float4 t = reinterpret_cast<const float4*>(x + row * 128)[lane];
    if (res) {
        const float4 r = reinterpret_cast<const float4*>(res + row * 128)[lane];
        t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
    }
    const float mean = warp_sum((t.x + t.y) + (t.z + t.w)) * (1.f / 128.f);
    const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
    const float rstd = rsqrtf(warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f) + LN_EPS);
    const float4 gm = reinterpret_cast<const float4*>(gamma)[lane];
    const float4 bt = reinterpret_cast<const float4*>(beta)[lane];
    float y[4] = {a * rstd * gm.x + bt.x, b * rstd * gm.y + bt.y, c * rstd * gm.z + bt.z, d * rstd * gm.w + bt.w};
    if (relu) { y[0] = fmaxf(y[0], 0.f); y[1] = fmaxf(y[1], 0.f); y[2] = fmaxf(y[2], 0.f); y[3] = fmaxf(y[3], 0.f); }
    if (out32) reinterpret_cast<float4*>(out32 + row * 128)[lane] = make_float4(y[0], y[1], y[2], y[3]);
    if (hi) {
        const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(y[0] - f01.x, y[1] - f01.y), l23 = __floats2half2_rn(y[2] - f23.x, y[3] - f23.y);
        uint2 uh, ul;
        uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
        ul.x = *reinterpret_cast<const uint32_t*>(&l01); ul.y = *reinterpret_cast<const uint32_t*>(&l23);
        reinterpret_cast<uint2*>(hi + row * 128)[lane] = uh;
        reinterpret_cast<uint2*>(lo + row * 128)[lane] = ul;
    }
}
void launch_layernorm_hl(const float* x, const float* res, const float* gamma, const float* beta, float* out32, __half* hi,
                         __half* lo, int64_t rows, int relu, cudaStream_t st) {
    if (rows <= 0) return;
    k_layernorm_hl<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, res, gamma, beta, out32, hi, lo, rows, relu);
    ++g_launches;
}

// fp32 -> fp16 (hi, lo) operand split
__global__ void k_split_hl(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
    ul.x = *reinterpret_cast<const uint32_t*>(&l01); ul.y = *reinterpret_cast<const uint32_t*>(&l23);
    reinterpret_cast<uint2*>(hi)[i] = uh;
    reinterpret_cast<uint2*>(lo)[i] = ul;
}
void launch_split_hl(const float* x, __half* hi, __half* lo, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const int64_t n4 = n / 4;
    k_split_hl<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, hi, lo, n4);
    ++g_launches;
}

__global__ void k_group_max(const float* __restrict__ in, float* __restrict__ out, int64_t G, int g, int D) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= G * D) return;
    const int64_t grp = idx / D;
    const int c = (int)(idx - grp * D);
    float m = -INFINITY;
    for (int r = 0; r < g; ++r) m = fmaxf(m, in[(grp * g + r) * D + c]);
    out[idx] = m;
}
void launch_group_max(const float* in, float* out, int64_t G, int g, int D, cudaStream_t st) {
    if (G <= 0) return;
    const int64_t n = G * D;
    k_group_max<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, G, g, D);
    ++g_launches;
}

// ------------------------------------------------------------------------------------------
// ActorNet: 4-scale 1-D ResNet FPN, GroupNorm(1 group) (network.py:12-61, layers.py:36-60,140-188)
// one CTA (256 threads) per actor; every activation lives in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int AN_THREADS = 256;
constexpr int AN_MAXLEN = 24;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < AN_THREADS / 32; ++i) t += red[i];
    return t;
}

// out[co][t] = sum_ci sum_k wT[(ci*3+k)*Cout+co] * in[ci][t*stride+k-1]   (zero padded)
__device__ void an_conv3(const float* __restrict__ in, int Cin, int Lin, const float* __restrict__ wT, int Cout,
                         int stride, float* __restrict__ out, int Lout) {
    const int nt = AN_THREADS / Cout;               // >= 1 (Cout <= 256)
    const int len = (Lout + nt - 1) / nt;           // <= AN_MAXLEN
    const int co = threadIdx.x % Cout;
    const int t0 = (threadIdx.x / Cout) * len;
    float acc[AN_MAXLEN];
#pragma unroll
    for (int t = 0; t < AN_MAXLEN; ++t) acc[t] = 0.f;
    if (threadIdx.x < nt * Cout) {
        for (int ci = 0; ci < Cin; ++ci) {
            const float w0 = __ldg(wT + (ci * 3 + 0) * Cout + co);
            const float w1 = __ldg(wT + (ci * 3 + 1) * Cout + co);
            const float w2 = __ldg(wT + (ci * 3 + 2) * Cout + co);
            const float* x = in + ci * Lin;
#pragma unroll
            for (int t = 0; t < AN_MAXLEN; ++t) {
                const int tt = t0 + t;
                if (t < len && tt < Lout) {
                    const int p = tt * stride;
                    const float xm = (p - 1 >= 0) ? x[p - 1] : 0.f;
                    const float x0 = x[p];
                    const float xp = (p + 1 < Lin) ? x[p + 1] : 0.f;
                    acc[t] = fmaf(w2, xp, fmaf(w1, x0, fmaf(w0, xm, acc[t])));
                }
            }
        }
#pragma unroll
        for (int t = 0; t < AN_MAXLEN; ++t) {
            const int tt = t0 + t;
            if (t < len && tt < Lout) out[co * Lout + tt] = acc[t];
        }
    }
    __syncthreads();
}

// 1x1 strided conv: out[co][t] = sum_ci wT[ci*Cout+co] * in[ci][t*stride]
__device__ void an_conv1(const float* __restrict__ in, int Cin, int Lin, const float* __restrict__ wT, int Cout,
                         int stride, float* __restrict__ out, int Lout) {
    for (int idx = threadIdx.x; idx < Cout * Lout; idx += AN_THREADS) {
        const int t = idx / Cout, co = idx - t * Cout;
        float a = 0.f;
        for (int ci = 0; ci < Cin; ++ci) a = fmaf(__ldg(wT + ci * Cout + co), in[ci * Lin + t * stride], a);
        out[co * Lout + t] = a;
    }
    __syncthreads();
}

// in-place GroupNorm(1 group) over C*L values, affine per channel; optional (+add) and ReLU
__device__ void an_gn(float* __restrict__ x, int C, int L, const float* __restrict__ gw, const float* __restrict__ gb,
                      const float* __restrict__ add, bool relu, float* red) {
    const int n = C * L;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += AN_THREADS) s += x[i];
    const float mean = block_sum(s, red) / (float)n;
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += AN_THREADS) { const float d = x[i] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q, red) / (float)n + LN_EPS);
    for (int i = threadIdx.x; i < n; i += AN_THREADS) {
        const int c = i / L;
        float v = (x[i] - mean) * rstd * __ldg(gw + c) + __ldg(gb + c);
        if (add) v += add[i];
        if (relu) v = fmaxf(v, 0.f);
        x[i] = v;
    }
    __syncthreads();
}

struct AnRes { const float *c1, *c2, *b1w, *b1b, *b2w, *b2b, *ds, *dsw, *dsb; };

// Res1d (layers.py:140-188): out = ReLU(GN(conv2(ReLU(GN(conv1(x))))) + shortcut(x))
__device__ void an_res1d(const float* x, int Cin, int Lin, const AnRes& w, int Cout, int stride, float* h, float* sc,
                         float* out, float* red) {
    const int Lout = (Lin - 1) / stride + 1;
    an_conv3(x, Cin, Lin, w.c1, Cout, stride, h, Lout);
    an_gn(h, Cout, Lout, w.b1w, w.b1b, nullptr, true, red);
    an_conv3(h, Cout, Lout, w.c2, Cout, 1, out, Lout);
    const float* shortcut = x;
    if (w.ds) {
        an_conv1(x, Cin, Lin, w.ds, Cout, stride, sc, Lout);
        an_gn(sc, Cout, Lout, w.dsw, w.dsb, nullptr, false, red);
        shortcut = sc;
    }
    an_gn(out, Cout, Lout, w.b2w, w.b2b, shortcut, true, red);
}

constexpr int AN_SMALL = 1536;   // 32x48 = 64x24 = 128x12 = 256x6
constexpr int AN_BIG = 6144;     // 128x48
constexpr int AN_SMEM_FLOATS = 3 * AN_SMALL + 4 * AN_SMALL + 3 * AN_BIG + 32;

__global__ void __launch_bounds__(AN_THREADS, 1) k_actor_net(const float* __restrict__ actors, float* __restrict__ outp,
                                                             int n_actors, ActorNetWeights W) {
    extern __shared__ __align__(16) float sm[];
    float* bufA = sm;                       // ping
    float* bufH = bufA + AN_SMALL;          // conv1 output
    float* bufS = bufH + AN_SMALL;          // shortcut
    float* grp = bufS + AN_SMALL;           // 4 saved group outputs, AN_SMALL each
    float* big0 = grp + 4 * AN_SMALL;       // pyramid
    float* big1 = big0 + AN_BIG;
    float* big2 = big1 + AN_BIG;
    float* red = big2 + AN_BIG;
    const int a = blockIdx.x;
    if (a >= n_actors) return;
    for (int i = threadIdx.x; i < 14 * 48; i += AN_THREADS) bufA[i] = actors[(int64_t)a * 14 * 48 + i];
    __syncthreads();

    const int Cg[4] = {32, 64, 128, 256};
    const float* cur = bufA;
    int Cin = 14, L = 48;
    for (int g = 0; g < 4; ++g) {
        const int stride = (g == 0) ? 1 : 2;
        AnRes r0{W.g_conv1[g][0], W.g_conv2[g][0], W.g_bn1w[g][0], W.g_bn1b[g][0], W.g_bn2w[g][0], W.g_bn2b[g][0],
                 W.g_ds[g], W.g_dsw[g], W.g_dsb[g]};
        // block 0 -> big0 (used as scratch output), block 1 -> grp[g]
        an_res1d(cur, Cin, L, r0, Cg[g], stride, bufH, bufS, big0, red);
        L = (L - 1) / stride + 1;
        AnRes r1{W.g_conv1[g][1], W.g_conv2[g][1], W.g_bn1w[g][1], W.g_bn1b[g][1], W.g_bn2w[g][1], W.g_bn2b[g][1],
                 nullptr, nullptr, nullptr};
        an_res1d(big0, Cg[g], L, r1, Cg[g], 1, bufH, bufS, grp + g * AN_SMALL, red);
        cur = grp + g * AN_SMALL;
        Cin = Cg[g];
    }
    // FPN top-down (network.py:55-58): lateral = conv3 + GN (no act)
    const int Ls[4] = {48, 24, 12, 6};
    float* pyr = big0;     // running pyramid level
    float* lat = big1;
    float* tmp = big2;
    an_conv3(grp + 3 * AN_SMALL, 256, 6, W.lat_conv[3], 128, 1, pyr, 6);
    an_gn(pyr, 128, 6, W.lat_w[3], W.lat_b[3], nullptr, false, red);
    for (int i = 2; i >= 0; --i) {
        const int Lc = Ls[i], Lp = Ls[i + 1];
        an_conv3(grp + i * AN_SMALL, Cg[i], Lc, W.lat_conv[i], 128, 1, lat, Lc);
        an_gn(lat, 128, Lc, W.lat_w[i], W.lat_b[i], nullptr, false, red);
        // F.interpolate(scale 2, linear, align_corners=False) of pyr [128, Lp] + lat -> tmp [128, Lc]
        for (int idx = threadIdx.x; idx < 128 * Lc; idx += AN_THREADS) {
            const int c = idx / Lc, t = idx - c * Lc;
            float src = ((float)t + 0.5f) * 0.5f - 0.5f;
            src = fmaxf(src, 0.f);
            const int i0 = (int)floorf(src);
            const int i1 = min(i0 + 1, Lp - 1);
            const float lam = src - (float)i0;
            tmp[idx] = pyr[c * Lp + i0] * (1.f - lam) + pyr[c * Lp + i1] * lam + lat[idx];
        }
        __syncthreads();
        float* sw = pyr; pyr = tmp; tmp = sw;
    }
    // output Res1d(128,128) on [128,48]; keep only the last step (network.py:60)
    AnRes ro{W.out_conv1, W.out_conv2, W.out_bn1w, W.out_bn1b, W.out_bn2w, W.out_bn2b, nullptr, nullptr, nullptr};
    an_res1d(pyr, 128, 48, ro, 128, 1, lat, nullptr, tmp, red);
    for (int c = threadIdx.x; c < 128; c += AN_THREADS) outp[(int64_t)a * 128 + c] = tmp[c * 48 + 47];
}

void launch_actor_net(const float* actors, float* out, int n_actors, const ActorNetWeights& w, cudaStream_t st) {
    if (n_actors <= 0) return;
    const size_t smem = AN_SMEM_FLOATS * sizeof(float);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_actor_net, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    k_actor_net<<<n_actors, AN_THREADS, smem, st>>>(actors, out, n_actors, w);
    ++g_launches;
}

// ------------------------------------------------------------------------------------------
// token scatter / gather  (network.py:320-324, 334-336)
// ------------------------------------------------------------------------------------------
__global__ void k_scatter_tokens(const float* __restrict__ ap, const float* __restrict__ lp,
                                 const SceneDesc* __restrict__ sd, float* __restrict__ x, int B, int Nmax) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)B * Nmax) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(row / Nmax), t = (int)(row - (int64_t)b * Nmax);
    const SceneDesc d = sd[b];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < d.n_actor) v = reinterpret_cast<const float4*>(ap + (int64_t)(d.actor_off + t) * 128)[lane];
    else if (t < d.n_actor + d.n_lane)
        v = reinterpret_cast<const float4*>(lp + (int64_t)(d.lane_off + t - d.n_actor) * 128)[lane];
    reinterpret_cast<float4*>(x + row * 128)[lane] = v;
}
void launch_scatter_tokens(const float* ap, const float* lp, const SceneDesc* sd, float* x, int B, int Nmax,
                           cudaStream_t st) {
    const int64_t rows = (int64_t)B * Nmax;
    k_scatter_tokens<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(ap, lp, sd, x, B, Nmax);
    ++g_launches;
}

__global__ void k_gather_tokens(const float* __restrict__ x, const SceneDesc* __restrict__ sd, float* __restrict__ actors,
                                float* __restrict__ cls, int B, int Nmax) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)B * Nmax) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(row / Nmax), t = (int)(row - (int64_t)b * Nmax);
    const SceneDesc d = sd[b];
    const float4 v = reinterpret_cast<const float4*>(x + row * 128)[lane];
    if (t < d.n_actor) reinterpret_cast<float4*>(actors + (int64_t)(d.actor_off + t) * 128)[lane] = v;
    else if (t == d.n_actor + d.n_lane) reinterpret_cast<float4*>(cls + (int64_t)b * 128)[lane] = v;
}
void launch_gather_tokens(const float* x, const SceneDesc* sd, float* actors, float* cls, int B, int Nmax,
                          cudaStream_t st) {
    const int64_t rows = (int64_t)B * Nmax;
    k_gather_tokens<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, sd, actors, cls, B, Nmax);
    ++g_launches;
}

// ------------------------------------------------------------------------------------------
// edge init: proj_rpe_scene (5 -> 128, LN, ReLU) over the [M,M] pair grid, zero cls row / col
// (network.py:326-330; get_rpe utils.py:193-242 when evaluated from anchors)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__half* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_edge_init(const SceneDesc* __restrict__ sd, const float* __restrict__ ctrs,
                                                   const float* __restrict__ vecs, const float* __restrict__ W,
                                                   const float* __restrict__ bias, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, OutT* __restrict__ edge, int b0,
                                                   int nb, int Nmax) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t per = (int64_t)Nmax * Nmax;
    if (row >= per * nb) return;
    const int bl = (int)(row / per);
    const int64_t rem = row - (int64_t)bl * per;
    const int i = (int)(rem / Nmax), j = (int)(rem - (int64_t)i * Nmax);
    const SceneDesc d = sd[b0 + bl];
    const int M = d.n_actor + d.n_lane;
    OutT* dst = edge + row * 128 + lane * 4;
    if (i >= M || j >= M) { store4(dst, make_float4(0.f, 0.f, 0.f, 0.f)); return; }
    float r[5];
    if (d.rpe) {
#pragma unroll
        for (int k = 0; k < 5; ++k) r[k] = __ldg(d.rpe + ((int64_t)k * M + i) * M + j);
    } else {
        // entry [i, j]: v1 = vecs[j], v2 = vecs[i], dpos = ctrs[j] - ctrs[i]   (utils.py:195-209)
        const float2 ci = reinterpret_cast<const float2*>(ctrs)[d.geom_off + i];
        const float2 cj = reinterpret_cast<const float2*>(ctrs)[d.geom_off + j];
        const float2 vi = reinterpret_cast<const float2*>(vecs)[d.geom_off + i];
        const float2 vj = reinterpret_cast<const float2*>(vecs)[d.geom_off + j];
        const float dx = cj.x - ci.x, dy = cj.y - ci.y;
        const float dist = sqrtf(dx * dx + dy * dy);
        const float nj = sqrtf(vj.x * vj.x + vj.y * vj.y), ni = sqrtf(vi.x * vi.x + vi.y * vi.y);
        const float den1 = nj * ni + 1e-10f, den2 = nj * dist + 1e-10f;
        r[0] = (vj.x * vi.x + vj.y * vi.y) / den1;
        r[1] = (vj.x * vi.y - vj.y * vi.x) / den1;
        r[2] = (vj.x * dx + vj.y * dy) / den2;
        r[3] = (vj.x * dy - vj.y * dx) / den2;
        r[4] = dist * 2.f / 100.f;
    }
    float y[4];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = lane * 4 + e;
        float a = __ldg(bias + c);
#pragma unroll
        for (int k = 0; k < 5; ++k) a = fmaf(__ldg(W + c * 5 + k), r[k], a);
        y[e] = a;
        s += a;
    }
    const float mean = warp_sum(s) * (1.f / 128.f);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float t = y[e] - mean; q += t * t; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / 128.f) + LN_EPS);
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = lane * 4 + e;
        o[e] = fmaxf((y[e] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c), 0.f);
    }
    store4(dst, make_float4(o[0], o[1], o[2], o[3]));
}

void launch_edge_init_f32(const SceneDesc* sd, const float* ctrs, const float* vecs, const float* W, const float* b,
                          const float* g, const float* be, float* edge, int b0, int nb, int Nmax, cudaStream_t st) {
    const int64_t rows = (int64_t)nb * Nmax * Nmax;
    if (rows <= 0) return;
    k_edge_init<float><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(sd, ctrs, vecs, W, b, g, be, edge, b0, nb, Nmax);
    ++g_launches;
}
// fp16 edge stream for the tensor-core path: packed fp32 pair helpers, then the kernel
namespace {
typedef unsigned long long ef2;
__device__ __forceinline__ ef2 e_pk2(float a, float b) { ef2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void e_upk2(ef2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ ef2 e_fma2(ef2 a, ef2 b, ef2 c) { ef2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ ef2 e_add2(ef2 a, ef2 b) { ef2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ ef2 e_sub2(ef2 a, ef2 b) { ef2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ ef2 e_mul2(ef2 a, ef2 b) { ef2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t e_cvt_relu_h2(ef2 v) {      // (lo, hi) fp32 -> fp16 pair, ReLU folded into the conversion
    float a, b; e_upk2(v, a, b); uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a)); return d;
}
}  // namespace

// ---- edge init, channel-parallel form ----------------------------------------------------------------------------
// LayerNorm of a 5 -> 128 projection has closed-form statistics: with Wc = W - mean_c(W), bc = b - mean_c(b) (centred over
// the 128 channels) x[c] - mean = Wc[c].r + bc[c], so var(r) = r^T Q r + L.r + c0 with Q = Wc^T Wc / 128, L = 2 Wc^T bc / 128,
// c0 = bc.bc / 128 (21 coefficients, formed in double on the host; centred, so no cancellation), and
//     y[c] = G[c].(rstd r) + g5[c] rstd + beta[c],   G = gamma Wc,  g5 = gamma bc.
// Nothing has to be held per channel across a reduction, so the work is laid out the other way round from a row per
// thread (the previous kernel: 0.70 ms, bound by 256 broadcast parameter loads per row): lane l owns channels [4l, 4l+4) with its 28 parameters in REGISTERS for the whole kernel (no
// parameter loads at all inside the loop), a warp takes 32 pair rows:
// phase 1, lane = row: RPE entry, variance, rstd -> (rstd r, rstd) through 64 B of shared memory per row; phase 2, lane =
// channels: 12 FFMA2 + 2 conversions per row and ONE coalesced 256-byte row store per warp instruction (no staging tile).
struct EiQuad { float q[21]; };     // c0, L[5], Q upper triangle by rows (off-diagonals doubled)
__global__ void __launch_bounds__(128) k_edge_init_ch(const SceneDesc* __restrict__ sd, const float* __restrict__ ctrs,
                                                      const float* __restrict__ vecs, const float2* __restrict__ tab,
                                                      const EiQuad quad, __half* __restrict__ edge, int nb, int Nmax, int min_tokens) {
    __shared__ __align__(16) float srow[4][32][16];          // per warp, per row: (r'_0, r'_0) ... (r'_4, r'_4), (rstd, rstd), valid, -
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ef2 P[2][7];                                             // [channel pair][G0..G4, g5, beta] of channels 4 lane + 2 pair + {0,1}
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
#pragma unroll
        for (int f = 0; f < 7; ++f) {
            const float2 v = __ldg(tab + (lane * 2 + pr) * 7 + f);
            P[pr][f] = e_pk2(v.x, v.y);
        }
    const int64_t per = (int64_t)Nmax * Nmax;
    const int groups_per_scene = (int)((per + 31) / 32);
    const int64_t n_groups = (int64_t)nb * groups_per_scene;
    const int64_t wstride = (int64_t)gridDim.x * 4;
    for (int64_t grp = (int64_t)blockIdx.x * 4 + warp; grp < n_groups; grp += wstride) {
        const int b = (int)(grp / groups_per_scene);
        const int64_t r0 = (grp - (int64_t)b * groups_per_scene) * 32;
        const SceneDesc d = sd[b];
        const int M = d.n_actor + d.n_lane;
        if (M + 1 < min_tokens) continue;                    // exact-tier scene (uniform per warp)
        {   // phase 1: this lane's row
            const int64_t r = r0 + lane;
            const int i = (int)(r / Nmax), j = (int)(r - (int64_t)i * Nmax);
            float rk[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            const bool ok = r < per && i < M && j < M;
            if (ok) {
                if (d.rpe) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) rk[k] = __ldg(d.rpe + ((int64_t)k * M + i) * M + j);
                } else {      // entry [i, j]: v1 = vecs[j], v2 = vecs[i], dpos = ctrs[j] - ctrs[i]   (utils.py:195-209)
                    const float2 ci = reinterpret_cast<const float2*>(ctrs)[d.geom_off + i];
                    const float2 cj = reinterpret_cast<const float2*>(ctrs)[d.geom_off + j];
                    const float2 vi = reinterpret_cast<const float2*>(vecs)[d.geom_off + i];
                    const float2 vj = reinterpret_cast<const float2*>(vecs)[d.geom_off + j];
                    const float dx = cj.x - ci.x, dy = cj.y - ci.y;
                    const float dist = sqrtf(dx * dx + dy * dy);
                    const float nj = sqrtf(vj.x * vj.x + vj.y * vj.y), ni = sqrtf(vi.x * vi.x + vi.y * vi.y);
                    const float den1 = nj * ni + 1e-10f, den2 = nj * dist + 1e-10f;
                    rk[0] = (vj.x * vi.x + vj.y * vi.y) / den1;
                    rk[1] = (vj.x * vi.y - vj.y * vi.x) / den1;
                    rk[2] = (vj.x * dx + vj.y * dy) / den2;
                    rk[3] = (vj.x * dy - vj.y * dx) / den2;
                    rk[4] = dist * 2.f / 100.f;
                }
            }
            float var = quad.q[0];
            int qi = 6;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                float t = quad.q[1 + k];
#pragma unroll
                for (int l = k; l < 5; ++l) t = fmaf(quad.q[qi++], rk[l], t);
                var = fmaf(rk[k], t, var);
            }
            const float rstd = rsqrtf(fmaxf(var, 0.f) + LN_EPS);
            float4* dst = reinterpret_cast<float4*>(&srow[warp][lane][0]);
            dst[0] = make_float4(rstd * rk[0], rstd * rk[0], rstd * rk[1], rstd * rk[1]);
            dst[1] = make_float4(rstd * rk[2], rstd * rk[2], rstd * rk[3], rstd * rk[3]);
            dst[2] = make_float4(rstd * rk[4], rstd * rk[4], rstd, rstd);
            dst[3] = make_float4(ok ? 1.f : 0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        // phase 2: lane = channels [4 lane, 4 lane + 4); one row per iteration, 8 bytes per lane = one 256-byte row per warp
        __half* dst0 = edge + ((int64_t)b * per + r0) * 128 + lane * 4;
        const int n_rows = (int)min((int64_t)32, per - r0);
#pragma unroll 4
        for (int rr = 0; rr < n_rows; ++rr) {
            const ulonglong2 A = *reinterpret_cast<const ulonglong2*>(&srow[warp][rr][0]);     // r'0, r'1
            const ulonglong2 Bv = *reinterpret_cast<const ulonglong2*>(&srow[warp][rr][4]);    // r'2, r'3
            const ulonglong2 Cv = *reinterpret_cast<const ulonglong2*>(&srow[warp][rr][8]);    // r'4, rstd
            const float okf = srow[warp][rr][12];
            uint32_t o[2];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                ef2 y = e_fma2(P[pr][5], Cv.y, P[pr][6]);
                y = e_fma2(P[pr][0], A.x, y);
                y = e_fma2(P[pr][1], A.y, y);
                y = e_fma2(P[pr][2], Bv.x, y);
                y = e_fma2(P[pr][3], Bv.y, y);
                y = e_fma2(P[pr][4], Cv.x, y);
                o[pr] = e_cvt_relu_h2(y);
            }
            if (okf == 0.f) { o[0] = 0u; o[1] = 0u; }         // padding rows of the [Nmax x Nmax] grid are zero (uniform per row)
            *reinterpret_cast<uint2*>(dst0 + (int64_t)rr * 128) = make_uint2(o[0], o[1]);
        }
        __syncwarp();                                         // the row table is rewritten by the next group
    }
}

// host side of k_edge_init_ch: lane parameter table [32 lanes][2 pairs][7] float2 and the 21 variance coefficients
void edge_init_pack_ch(const float* W, const float* b, const float* g, const float* be, float* tab896, float* quad21) {
    double wm[5] = {0, 0, 0, 0, 0}, bm = 0;
    for (int c = 0; c < 128; ++c) { for (int k = 0; k < 5; ++k) wm[k] += W[c * 5 + k]; bm += b[c]; }
    for (int k = 0; k < 5; ++k) wm[k] /= 128.0;
    bm /= 128.0;
    double Q[5][5] = {}, Lv[5] = {}, c0 = 0;
    for (int c = 0; c < 128; ++c) {
        double wc[5];
        for (int k = 0; k < 5; ++k) wc[k] = W[c * 5 + k] - wm[k];
        const double bc = b[c] - bm;
        for (int k = 0; k < 5; ++k) { Lv[k] += 2.0 * wc[k] * bc; for (int l = 0; l < 5; ++l) Q[k][l] += wc[k] * wc[l]; }
        c0 += bc * bc;
        const int lane = c >> 2, pr = (c >> 1) & 1, e = c & 1;
        float* t = tab896 + ((lane * 2 + pr) * 7) * 2 + e;     // float2 entries: [..][f] = (channel 2p, channel 2p + 1)
        for (int k = 0; k < 5; ++k) t[k * 2] = (float)(g[c] * wc[k]);
        t[5 * 2] = (float)(g[c] * bc);
        t[6 * 2] = be[c];
    }
    quad21[0] = (float)(c0 / 128.0);
    for (int k = 0; k < 5; ++k) quad21[1 + k] = (float)(Lv[k] / 128.0);
    int qi = 6;
    for (int k = 0; k < 5; ++k)
        for (int l = k; l < 5; ++l) quad21[qi++] = (float)((l == k ? 1.0 : 2.0) * Q[k][l] / 128.0);
}

void launch_edge_init_ch(const SceneDesc* sd, const float* ctrs, const float* vecs, const float* dev_tab, const float* quad21,
                         __half* edge, int b0, int nb, int Nmax, int min_tokens, cudaStream_t st) {
    const int64_t groups = (int64_t)nb * (((int64_t)Nmax * Nmax + 31) / 32);
    if (groups <= 0) return;
    EiQuad q;
    for (int i = 0; i < 21; ++i) q.q[i] = quad21[i];
    const int64_t blocks = std::min<int64_t>((groups + 3) / 4, 148 * 16);
    k_edge_init_ch<<<(unsigned)blocks, 128, 0, st>>>(sd + b0, ctrs, vecs, reinterpret_cast<const float2*>(dev_tab), q, edge, nb, Nmax, min_tokens);
    ++g_launches;
}


// ------------------------------------------------------------------------------------------
// exact-path pair epilogues (network.py:197-202, 222)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ln4(float4 v, const float* gamma, const float* beta, int lane, bool relu) {
    const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    const float rstd = rsqrtf(warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f) + LN_EPS);
    const float4 gm = reinterpret_cast<const float4*>(gamma)[lane];
    const float4 bt = reinterpret_cast<const float4*>(beta)[lane];
    float4 o = make_float4(a * rstd * gm.x + bt.x, b * rstd * gm.y + bt.y, c * rstd * gm.z + bt.z, d * rstd * gm.w + bt.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    return o;
}

__global__ void __launch_bounds__(256) k_pair_memory_epi(const float* __restrict__ tmp, const float* __restrict__ stq,
                                                         const float* __restrict__ g, const float* __restrict__ be,
                                                         float* __restrict__ memory, int b0, int nb, int Nmax) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t per = (int64_t)Nmax * Nmax;
    if (row >= per * nb) return;
    const int bl = (int)(row / per);
    const int64_t rem = row - (int64_t)bl * per;
    const int i = (int)(rem / Nmax), j = (int)(rem - (int64_t)i * Nmax);
    const int64_t tb = (int64_t)(b0 + bl) * Nmax;
    float4 v = reinterpret_cast<const float4*>(tmp + row * 128)[lane];
    const float4 s = reinterpret_cast<const float4*>(stq + (tb + j) * 384)[lane];          // S[j]  (src_x = node[j])
    const float4 t = reinterpret_cast<const float4*>(stq + (tb + i) * 384 + 128)[lane];    // T[i]  (tar_x = node[i])
    v.x = (v.x + s.x) + t.x; v.y = (v.y + s.y) + t.y; v.z = (v.z + s.z) + t.z; v.w = (v.w + s.w) + t.w;
    reinterpret_cast<float4*>(memory + row * 128)[lane] = ln4(v, g, be, lane, true);
}
void launch_pair_memory_epi(const float* tmp, const float* stq, const float* g, const float* be, float* memory, int b0,
                            int nb, int Nmax, cudaStream_t st) {
    const int64_t rows = (int64_t)nb * Nmax * Nmax;
    if (rows <= 0) return;
    k_pair_memory_epi<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(tmp, stq, g, be, memory, b0, nb, Nmax);
    ++g_launches;
}

__global__ void __launch_bounds__(256) k_pair_edge_epi(const float* __restrict__ tmp, const float* __restrict__ gp,
                                                       const float* __restrict__ bp, const float* __restrict__ ge,
                                                       const float* __restrict__ bee, float* __restrict__ edge,
                                                       int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4 u = ln4(reinterpret_cast<const float4*>(tmp + row * 128)[lane], gp, bp, lane, true);
    float4 e = reinterpret_cast<const float4*>(edge + row * 128)[lane];
    e.x += u.x; e.y += u.y; e.z += u.z; e.w += u.w;
    reinterpret_cast<float4*>(edge + row * 128)[lane] = ln4(e, ge, bee, lane, false);
}
void launch_pair_edge_epi(const float* tmp, const float* gp, const float* bp, const float* ge, const float* bee,
                          float* edge, int64_t rows, cudaStream_t st) {
    if (rows <= 0) return;
    k_pair_edge_epi<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(tmp, gp, bp, ge, bee, edge, rows);
    ++g_launches;
}

// one CTA (128 threads = channels) per (scene, query j); keys i = 0..N_b-1 (no mask: network.py:332)
__global__ void __launch_bounds__(128) k_pair_attention(const float* __restrict__ kv, const float* __restrict__ stq,
                                                        const SceneDesc* __restrict__ sd, float* __restrict__ attn,
                                                        int b0, int nb, int Nmax) {
    extern __shared__ float sc[];   // [Nmax][8]
    const int bl = blockIdx.x / Nmax, j = blockIdx.x - bl * Nmax;
    const int b = b0 + bl;
    const SceneDesc d = sd[b];
    const int N = d.n_actor + d.n_lane + 1;
    const int c = threadIdx.x;
    float* dst = attn + ((int64_t)b * Nmax + j) * 128 + c;
    if (j >= N) { *dst = 0.f; return; }
    const float q = stq[((int64_t)b * Nmax + j) * 384 + 256 + c];   // already scaled by 1/sqrt(16)
    const int64_t base = ((int64_t)bl * Nmax) * Nmax + j;
    for (int i = 0; i < N; ++i) {
        float p = q * kv[(base + (int64_t)i * Nmax) * 256 + c];
        p += __shfl_xor_sync(0xffffffffu, p, 8);
        p += __shfl_xor_sync(0xffffffffu, p, 4);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        if ((c & 15) == 0) sc[i * 8 + (c >> 4)] = p;
    }
    __syncthreads();
    if (c < 8) {
        float m = -INFINITY;
        for (int i = 0; i < N; ++i) m = fmaxf(m, sc[i * 8 + c]);
        float s = 0.f;
        for (int i = 0; i < N; ++i) { const float e = expf(sc[i * 8 + c] - m); sc[i * 8 + c] = e; s += e; }
        const float inv = 1.f / s;
        for (int i = 0; i < N; ++i) sc[i * 8 + c] *= inv;
    }
    __syncthreads();
    float o = 0.f;
    const int h = c >> 4;
    for (int i = 0; i < N; ++i) o = fmaf(sc[i * 8 + h], kv[(base + (int64_t)i * Nmax) * 256 + 128 + c], o);
    *dst = o;
}
void launch_pair_attention(const float* kv, const float* stq, const SceneDesc* sd, float* attn, int b0, int nb, int Nmax,
                           cudaStream_t st) {
    if (nb <= 0) return;
    k_pair_attention<<<nb * Nmax, 128, (size_t)Nmax * 8 * sizeof(float), st>>>(kv, stq, sd, attn, b0, nb, Nmax);
    ++g_launches;
}

// ------------------------------------------------------------------------------------------
// decoder pieces
// ------------------------------------------------------------------------------------------
// nn.TransformerEncoderLayer self-attention over the 6 modes (network.py:378-380,502): 4 heads x 32
__global__ void __launch_bounds__(128) k_mode_attention(const float* __restrict__ qkv, float* __restrict__ out, int B) {
    __shared__ float s_qkv[6][384];
    __shared__ float s_p[4][6][6];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < 6 * 384; i += 128) s_qkv[i / 384][i % 384] = qkv[(int64_t)b * 6 * 384 + i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * 36; idx += 128) {
        const int h = idx / 36, r = idx % 36, s = r / 6, t = r % 6;
        float a = 0.f;
        for (int dd = 0; dd < 32; ++dd) a = fmaf(s_qkv[s][h * 32 + dd], s_qkv[t][128 + h * 32 + dd], a);
        s_p[h][s][t] = a * 0.17677669529663687f;   // 1/sqrt(32)
    }
    __syncthreads();
    if (threadIdx.x < 24) {
        const int h = threadIdx.x / 6, s = threadIdx.x % 6;
        float m = -INFINITY;
        for (int t = 0; t < 6; ++t) m = fmaxf(m, s_p[h][s][t]);
        float sum = 0.f;
        for (int t = 0; t < 6; ++t) { const float e = expf(s_p[h][s][t] - m); s_p[h][s][t] = e; sum += e; }
        for (int t = 0; t < 6; ++t) s_p[h][s][t] /= sum;
    }
    __syncthreads();
    const int c = threadIdx.x, h = c >> 5;
    for (int s = 0; s < 6; ++s) {
        float o = 0.f;
        for (int t = 0; t < 6; ++t) o = fmaf(s_p[h][s][t], s_qkv[t][256 + c], o);
        out[((int64_t)b * 6 + s) * 128 + c] = o;
    }
}
void launch_mode_attention(const float* qkv, float* out, int B, cudaStream_t st) {
    if (B <= 0) return;
    k_mode_attention<<<B, 128, 0, st>>>(qkv, out, B);
    ++g_launches;
}

__global__ void k_embed_combine(const float* __restrict__ ce, const float* __restrict__ ae, const float* __restrict__ tgt,
                                const int32_t* __restrict__ actor_scene, float* __restrict__ embed, int n_actors) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n_actors*6*128
    if (idx >= (int64_t)n_actors * 768) return;
    const int a = (int)(idx / 768);
    const int r = (int)(idx - (int64_t)a * 768);
    const int m = r >> 7, c = r & 127;
    const int b = actor_scene[a];
    float v = ce[((int64_t)b * 6 + m) * 128 + c] + ae[idx];
    if (m == 0) v += tgt[(int64_t)b * 128 + c];     // quirk: mode 0 of every actor (network.py:506-508)
    embed[idx] = v;
}
void launch_embed_combine(const float* ce, const float* ae, const float* tgt, const int32_t* actor_scene, float* embed,
                          int n_actors, cudaStream_t st) {
    if (n_actors <= 0) return;
    const int64_t n = (int64_t)n_actors * 768;
    k_embed_combine<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ce, ae, tgt, actor_scene, embed, n_actors);
    ++g_launches;
}

__global__ void k_softmax6(const float* __restrict__ logits, float* __restrict__ cls, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float v[6], m = -INFINITY, s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { v[i] = logits[b * 6 + i]; m = fmaxf(m, v[i]); }
#pragma unroll
    for (int i = 0; i < 6; ++i) { v[i] = expf(v[i] - m); s += v[i]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) cls[b * 6 + i] = v[i] / s;
}
void launch_softmax6(const float* logits, float* cls, int B, cudaStream_t st) {
    if (B <= 0) return;
    k_softmax6<<<(B + 127) / 128, 128, 0, st>>>(logits, cls, B);
    ++g_launches;
}

// Bezier decode (network.py:515-523,545): param row [8,5] -> reg [60,5], vel [60,2], cov_vel [60,3]
__global__ void __launch_bounds__(64) k_bezier(const float* __restrict__ param, const float* __restrict__ T,
                                               const float* __restrict__ Tp, float* __restrict__ reg,
                                               float* __restrict__ vel, float* __restrict__ cov_vel, int n_rows) {
    __shared__ float P[40];
    const int row = blockIdx.x;
    if (threadIdx.x < 40) P[threadIdx.x] = param[(int64_t)row * 40 + threadIdx.x];
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= 60) return;
    float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float w = __ldg(T + t * 8 + i);
#pragma unroll
        for (int d = 0; d < 5; ++d) r[d] = fmaf(w, P[i * 5 + d], r[d]);
    }
    float dv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float w = __ldg(Tp + t * 7 + i);
#pragma unroll
        for (int d = 0; d < 5; ++d) dv[d] = fmaf(w, P[(i + 1) * 5 + d] - P[i * 5 + d], dv[d]);
    }
    float* ro = reg + ((int64_t)row * 60 + t) * 5;
    ro[0] = r[0]; ro[1] = r[1]; ro[2] = expf(r[2]); ro[3] = expf(r[3]); ro[4] = expf(r[4]);
    float* vo = vel + ((int64_t)row * 60 + t) * 2;
    vo[0] = dv[0] / 6.0f; vo[1] = dv[1] / 6.0f;
    if (cov_vel) {
        float* co = cov_vel + ((int64_t)row * 60 + t) * 3;
        co[0] = dv[2] / 6.0f; co[1] = dv[3] / 6.0f; co[2] = dv[4] / 6.0f;
    }
}
void launch_bezier(const float* param, const float* T, const float* Tp, float* reg, float* vel, float* cov_vel,
                   int n_rows, cudaStream_t st) {
    if (n_rows <= 0) return;
    k_bezier<<<n_rows, 64, 0, st>>>(param, T, Tp, reg, vel, cov_vel, n_rows);
    ++g_launches;
}

}  // namespace mind
