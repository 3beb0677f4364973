// Internal launch interface between mind_api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace mind {

// running launch counter (bench.py gpu_launches)
extern int64_t g_launches;

// Pinned host staging for the small per-forward descriptor tables.  cudaMemcpyAsync from pageable memory stalls the
// host until the stream reaches the copy once a table exceeds the driver's staging threshold (the 90 KB work list of
// a 256-scene batch did: the serving loop lost its two-deep pipelining); from this ring every upload is truly
// asynchronous.  One slot per in-flight forward; a slot is recycled after its event has completed.
struct HostStage {
    static constexpr int kSlots = 4;
    struct Slot { char* host = nullptr; size_t cap = 0, used = 0; cudaEvent_t done = nullptr; bool pending = false; };
    Slot slot[kSlots];
    int cur = 0;
    const char* begin();                                                        // next slot (waits only if 4 forwards are in flight)
    const char* upload(void* dev_dst, const void* src, size_t bytes, cudaStream_t st);
    void end(cudaStream_t st);                                                  // records the slot's event
    void release();
};

struct GemmArgs {
    const float* A = nullptr; int lda = 0;      // [M,K]
    const float* W = nullptr; int ldw = 0;      // [N,K] (torch Linear layout), row stride ldw
    const float* bias = nullptr;                // [N] or null
    const float* gbias = nullptr; int gsize = 1; int ldg = 0;   // per row-group bias [(m/gsize), N]
    float* C = nullptr; int ldc = 0;            // [M,N]
    int M = 0, N = 0, K = 0;
    int relu = 0;
    // set by launch_gemm for skinny outputs with a long K (few CTAs, serial K loop): grid.z slices of K write raw partial
    // sums to `part` [ksplit][M][N]; a second kernel adds them in slice order (deterministic) and applies bias / ReLU
    int ksplit = 1; float* part = nullptr;
};
void launch_gemm(const GemmArgs& g, cudaStream_t st);

// out[r,:] = act(LN(x[r,:] (+ res[r,:]))) ; D <= 1536, D % 4 == 0
void launch_layernorm(const float* x, const float* res, const float* gamma, const float* beta,
                      float* out, int64_t rows, int D, int relu, cudaStream_t st);

// D = 128 LayerNorm that also writes the fp16 (hi, lo) operand split for the tensor-core GEMM engine
void launch_layernorm_hl(const float* x, const float* res, const float* gamma, const float* beta, float* out32, __half* hi,
                         __half* lo, int64_t rows, int relu, cudaStream_t st);

// fp32 -> fp16 (hi, lo) split of n floats (n % 4 == 0)
void launch_split_hl(const float* x, __half* hi, __half* lo, int64_t n, cudaStream_t st);

// max over groups of `g` consecutive rows: in [G*g, D] -> out [G, D]
void launch_group_max(const float* in, float* out, int64_t G, int g, int D, cudaStream_t st);

// ---- ActorNet (reference network.py:12-61) : one CTA per actor, all layers in shared memory
struct ActorNetWeights {
    // conv filters pre-transposed to [ci][k][co]; 1x1 shortcut filters to [ci][co]
    const float* g_conv1[4][2]; const float* g_conv2[4][2];
    const float* g_bn1w[4][2]; const float* g_bn1b[4][2];
    const float* g_bn2w[4][2]; const float* g_bn2b[4][2];
    const float* g_ds[4]; const float* g_dsw[4]; const float* g_dsb[4];   // block 0 of each group
    const float* lat_conv[4]; const float* lat_w[4]; const float* lat_b[4];
    const float* out_conv1; const float* out_conv2;
    const float* out_bn1w; const float* out_bn1b; const float* out_bn2w; const float* out_bn2b;
};
void launch_actor_net(const float* actors, float* out, int n_actors, const ActorNetWeights& w, cudaStream_t st);

// ---- scene descriptor table (device) -------------------------------------------------------
struct SceneDesc {
    int32_t actor_off, lane_off;   // offsets into the compact actor / lane arrays
    int32_t n_actor, n_lane;       // N = n_actor + n_lane + 1 tokens
    int32_t geom_off;              // offset into ctrs/vecs (= actor_off + lane_off)
    int32_t pad_;
    const float* rpe;              // dev [5, M, M] or null
};

// tokens x [B, Nmax, 128]: scatter projected actors / lanes, zero cls + padding rows
void launch_scatter_tokens(const float* actors_p, const float* lanes_p, const SceneDesc* sd, float* x,
                           int B, int Nmax, cudaStream_t st);
// gather fused actor tokens -> [sumNa,128] and cls tokens -> [B,128]
void launch_gather_tokens(const float* x, const SceneDesc* sd, float* actors, float* cls, int B, int Nmax,
                          cudaStream_t st);

// edge0[b,i,j,:] = ReLU(LN(W5 . rpe[b,:,i,j] + b)) for i,j < M_b, zero elsewhere (network.py:326-330)
// OutT = float (exact path) or __half (tensor-core path)
void launch_edge_init_f32(const SceneDesc* sd, const float* ctrs, const float* vecs, const float* W, const float* b,
                          const float* g, const float* be, float* edge, int b0, int nb, int Nmax, cudaStream_t st);

// tensor-core mode (fp16 edge stream): closed-form LayerNorm statistics, channel-parallel (simt_kernels.cu, k_edge_init_ch);
// the table packs gamma * centred W / b and beta per lane, quad21 the variance coefficients
void edge_init_pack_ch(const float* W, const float* b, const float* g, const float* be, float* tab896, float* quad21);
void launch_edge_init_ch(const SceneDesc* sd, const float* ctrs, const float* vecs, const float* dev_tab, const float* quad21,
                         __half* edge, int b0, int nb, int Nmax, int min_tokens, cudaStream_t st);

// exact-path pair epilogues on rows r = ((b*Nmax + i)*Nmax + j)
// memory = ReLU(LN(tmp + S[b,j] + T[b,i]))   with STQ [B*Nmax, 384] = [S | T | q]
void launch_pair_memory_epi(const float* tmp, const float* stq, const float* g, const float* be, float* memory,
                            int b0, int nb, int Nmax, cudaStream_t st);
// edge = LN_e(edge + ReLU(LN_p(tmp)))
void launch_pair_edge_epi(const float* tmp, const float* gp, const float* bp, const float* ge, const float* bee,
                          float* edge, int64_t rows, cudaStream_t st);
// attn[b,j,:] = sum_i softmax_i(q[b,j,h].K[b,i,j,h]) V[b,i,j,h]   (KV [rows,256] = [K|V]; q pre-scaled)
void launch_pair_attention(const float* kv, const float* stq, const SceneDesc* sd, float* attn, int b0, int nb,
                           int Nmax, cudaStream_t st);

// ---- exact tier of the tensor-core mode (pair_x3.cu): row-wise epilogues between 3-term tcgen05 GEMMs on the compact
// pair grid row = (s*Ns + i)*Ns + j of the small scenes sids[0..ns); every kernel emits the next GEMM's (hi, lo) operand
void launch_x3_edge_init(const SceneDesc* sd, const int32_t* sids, const float* ctrs, const float* vecs, const float* W,
                         const float* b, const float* g, const float* be, float* e32, __half* eh, __half* el, int ns, int Ns,
                         cudaStream_t st);
void launch_x3_memory_epi(const float* tmp, const float* stq, const int32_t* sids, const float* g, const float* be, __half* mh,
                          __half* ml, int ns, int Ns, int Nmax, cudaStream_t st);
void launch_x3_edge_epi(const float* tmp, const float* gp, const float* bp, const float* ge, const float* bee, float* e32,
                        __half* eh, __half* el, int64_t rows, cudaStream_t st);
void launch_x3_attention(const float* kv, const float* stq, const SceneDesc* sd, const int32_t* sids, __half* ah, __half* al,
                         int ns, int Ns, int Nmax, cudaStream_t st);

// ---- decoder pieces (reference network.py:483-556) -----------------------------------------
// self attention over the 6 modes of each scene: qkv [B*6,384] -> out [B*6,128], 4 heads
void launch_mode_attention(const float* qkv, float* out, int B, cudaStream_t st);
// embed[a*6+m] = ce[scene(a)*6+m] + ae[a*6+m] + (m==0 ? tgt[scene(a)] : 0)
void launch_embed_combine(const float* ce, const float* ae, const float* tgt, const int32_t* actor_scene,
                          float* embed, int n_actors, cudaStream_t st);
void launch_softmax6(const float* logits, float* cls, int B, cudaStream_t st);
void launch_bezier(const float* param, const float* T, const float* Tp, float* reg, float* vel, float* cov_vel,
                   int n_rows /* sumNa*6 */, cudaStream_t st);

}  // namespace mind
