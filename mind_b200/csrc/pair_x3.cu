// Exact tier of the tensor-core precision mode ("x3"): the rela-fusion pair pipeline of scenes too small for the
// fp16-operand fused kernel (few keys per softmax: its operand rounding is not averaged out), un-fused, with every N^2
// contraction on the tcgen05 GEMM engine as a 3-term fp16 hi/lo product (fp32-equivalent) and the edge kept in fp32.
// These are the row-wise epilogues between the GEMMs; each also emits the (hi, lo) operand of the next GEMM.
// Pair rows live on a compact grid: row = (s * Ns + i) * Ns + j for the s-th small scene (scene id sids[s]), key i,
// query j.  Reference semantics: RelaFusionLayer, planners/mind/networks/network.py:182-232 and :326-330.
#include "kernels.h"
#include <math.h>

namespace mind {

#define LN_EPS 1e-5f

namespace {
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void store_hl(__half* hi, __half* lo, int64_t row, int lane, float4 y) {
    const __half2 h01 = __floats2half2_rn(y.x, y.y), h23 = __floats2half2_rn(y.z, y.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(y.x - f01.x, y.y - f01.y), l23 = __floats2half2_rn(y.z - f23.x, y.w - f23.y);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
    ul.x = *reinterpret_cast<const uint32_t*>(&l01); ul.y = *reinterpret_cast<const uint32_t*>(&l23);
    reinterpret_cast<uint2*>(hi + row * 128)[lane] = uh;
    reinterpret_cast<uint2*>(lo + row * 128)[lane] = ul;
}
__device__ __forceinline__ float4 ln4x(float4 v, const float* gamma, const float* beta, int lane, bool relu) {
    const float mean = wsum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    const float rstd = rsqrtf(wsum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f) + LN_EPS);
    const float4 gm = reinterpret_cast<const float4*>(gamma)[lane];
    const float4 bt = reinterpret_cast<const float4*>(beta)[lane];
    float4 o = make_float4(a * rstd * gm.x + bt.x, b * rstd * gm.y + bt.y, c * rstd * gm.z + bt.z, d * rstd * gm.w + bt.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    return o;
}

// edge0 = ReLU(LN(W5 . rpe + b)) on the [M, M] grid of each small scene, zero elsewhere (network.py:326-330;
// get_rpe utils.py:193-212 when evaluated from anchors): fp32 copy + (hi, lo) operand of layer 0's W_e product
__global__ void __launch_bounds__(256) k_x3_edge_init(const SceneDesc* __restrict__ sd, const int32_t* __restrict__ sids,
                                                      const float* __restrict__ ctrs, const float* __restrict__ vecs,
                                                      const float* __restrict__ W, const float* __restrict__ bias,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ e32, __half* __restrict__ eh, __half* __restrict__ el,
                                                      int ns, int Ns) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t per = (int64_t)Ns * Ns;
    if (row >= per * ns) return;
    const int s = (int)(row / per);
    const int64_t rem = row - (int64_t)s * per;
    const int i = (int)(rem / Ns), j = (int)(rem - (int64_t)i * Ns);
    const SceneDesc d = sd[sids[s]];
    const int M = d.n_actor + d.n_lane;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < M && j < M) {                       // warp-uniform: one pair row per warp
        float r[5];
        if (d.rpe) {
#pragma unroll
            for (int k = 0; k < 5; ++k) r[k] = __ldg(d.rpe + ((int64_t)k * M + i) * M + j);
        } else {
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
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = lane * 4 + e;
            float a = __ldg(bias + c);
#pragma unroll
            for (int k = 0; k < 5; ++k) a = fmaf(__ldg(W + c * 5 + k), r[k], a);
            y[e] = a;
        }
        o = ln4x(make_float4(y[0], y[1], y[2], y[3]), gamma, beta, lane, true);
    }
    reinterpret_cast<float4*>(e32 + row * 128)[lane] = o;
    store_hl(eh, el, row, lane, o);
}

// memory = ReLU(LN(tmp + S[j] + T[i])) -> (hi, lo)   (network.py:197-199; STQ [B*Nmax, 384] = S | T | q)
__global__ void __launch_bounds__(256) k_x3_memory_epi(const float* __restrict__ tmp, const float* __restrict__ stq,
                                                       const int32_t* __restrict__ sids, const float* __restrict__ g,
                                                       const float* __restrict__ be, __half* __restrict__ mh,
                                                       __half* __restrict__ ml, int ns, int Ns, int Nmax) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t per = (int64_t)Ns * Ns;
    if (row >= per * ns) return;
    const int s = (int)(row / per);
    const int64_t rem = row - (int64_t)s * per;
    const int i = (int)(rem / Ns), j = (int)(rem - (int64_t)i * Ns);
    const int64_t tb = (int64_t)sids[s] * Nmax;
    float4 v = reinterpret_cast<const float4*>(tmp + row * 128)[lane];
    const float4 sv = reinterpret_cast<const float4*>(stq + (tb + j) * 384)[lane];
    const float4 tv = reinterpret_cast<const float4*>(stq + (tb + i) * 384 + 128)[lane];
    v.x = (v.x + sv.x) + tv.x; v.y = (v.y + sv.y) + tv.y; v.z = (v.z + sv.z) + tv.z; v.w = (v.w + sv.w) + tv.w;
    store_hl(mh, ml, row, lane, ln4x(v, g, be, lane, true));
}

// edge = LN_e(edge + ReLU(LN_p(tmp)))  (network.py:202): fp32 in place + (hi, lo) operand of the next layer
__global__ void __launch_bounds__(256) k_x3_edge_epi(const float* __restrict__ tmp, const float* __restrict__ gp,
                                                     const float* __restrict__ bp, const float* __restrict__ ge,
                                                     const float* __restrict__ bee, float* __restrict__ e32,
                                                     __half* __restrict__ eh, __half* __restrict__ el, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4 u = ln4x(reinterpret_cast<const float4*>(tmp + row * 128)[lane], gp, bp, lane, true);
    float4 e = reinterpret_cast<const float4*>(e32 + row * 128)[lane];
    e.x += u.x; e.y += u.y; e.z += u.z; e.w += u.w;
    const float4 o = ln4x(e, ge, bee, lane, false);
    reinterpret_cast<float4*>(e32 + row * 128)[lane] = o;
    store_hl(eh, el, row, lane, o);
}

// attention of query j over the keys i < N of its scene with K = V-source = memory column j (network.py:222-225);
// one CTA (128 threads = channels) per (small scene, query).  Output = the (hi, lo) operand of the out-proj GEMM.
__global__ void __launch_bounds__(128) k_x3_attention(const float* __restrict__ kv, const float* __restrict__ stq,
                                                      const SceneDesc* __restrict__ sd, const int32_t* __restrict__ sids,
                                                      __half* __restrict__ ah, __half* __restrict__ al, int Ns, int Nmax) {
    extern __shared__ float sc[];   // [Ns][8]
    const int s = blockIdx.x / Ns, j = blockIdx.x - s * Ns;
    const int b = sids[s];
    const SceneDesc d = sd[b];
    const int N = d.n_actor + d.n_lane + 1;
    const int c = threadIdx.x;
    if (j >= N) return;                 // padded query rows stay zero (cleared once per forward)
    const float q = stq[((int64_t)b * Nmax + j) * 384 + 256 + c];   // already scaled by 1/sqrt(16)
    const int64_t base = ((int64_t)s * Ns) * Ns + j;
    for (int i = 0; i < N; ++i) {
        float p = q * kv[(base + (int64_t)i * Ns) * 256 + c];
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
        float sum = 0.f;
        for (int i = 0; i < N; ++i) { const float e = expf(sc[i * 8 + c] - m); sc[i * 8 + c] = e; sum += e; }
        const float inv = 1.f / sum;
        for (int i = 0; i < N; ++i) sc[i * 8 + c] *= inv;
    }
    __syncthreads();
    float o = 0.f;
    const int h = c >> 4;
    for (int i = 0; i < N; ++i) o = fmaf(sc[i * 8 + h], kv[(base + (int64_t)i * Ns) * 256 + 128 + c], o);
    const __half hh = __float2half_rn(o);
    const int64_t orow = ((int64_t)b * Nmax + j) * 128 + c;
    ah[orow] = hh;
    al[orow] = __float2half_rn(o - __half2float(hh));
}
}  // namespace

void launch_x3_edge_init(const SceneDesc* sd, const int32_t* sids, const float* ctrs, const float* vecs, const float* W,
                         const float* b, const float* g, const float* be, float* e32, __half* eh, __half* el, int ns, int Ns,
                         cudaStream_t st) {
    const int64_t rows = (int64_t)ns * Ns * Ns;
    if (rows <= 0) return;
    k_x3_edge_init<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(sd, sids, ctrs, vecs, W, b, g, be, e32, eh, el, ns, Ns);
    ++g_launches;
}
void launch_x3_memory_epi(const float* tmp, const float* stq, const int32_t* sids, const float* g, const float* be, __half* mh,
                          __half* ml, int ns, int Ns, int Nmax, cudaStream_t st) {
    const int64_t rows = (int64_t)ns * Ns * Ns;
    if (rows <= 0) return;
    k_x3_memory_epi<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(tmp, stq, sids, g, be, mh, ml, ns, Ns, Nmax);
    ++g_launches;
}
void launch_x3_edge_epi(const float* tmp, const float* gp, const float* bp, const float* ge, const float* bee, float* e32,
                        __half* eh, __half* el, int64_t rows, cudaStream_t st) {
    if (rows <= 0) return;
    k_x3_edge_epi<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(tmp, gp, bp, ge, bee, e32, eh, el, rows);
    ++g_launches;
}
void launch_x3_attention(const float* kv, const float* stq, const SceneDesc* sd, const int32_t* sids, __half* ah, __half* al,
                         int ns, int Ns, int Nmax, cudaStream_t st) {
    if (ns <= 0) return;
    k_x3_attention<<<(unsigned)(ns * Ns), 128, (size_t)Ns * 8 * sizeof(float), st>>>(kv, stq, sd, sids, ah, al, Ns, Nmax);
    ++g_launches;
}

}  // namespace mind
