// Tensor-core GEMM engine + ActorNet on top of it (see tc_gemm.cu / actor_tc.cu).
#pragma once
#include "kernels.h"
#include <map>
#include <string>
#include <vector>

namespace mind {

struct TcGemm {
    const void* amap_hi = nullptr;   // CUtensorMap (128 B) over the fp16 hi part of A
    const void* amap_lo = nullptr;   //   "   lo part (only when split)
    const void* wmap = nullptr;      // CUtensorMap over W [N][k_total] (k_total = 2*k_blocks*64 when split: [hi | lo])
    int split = 1;                   // 3-term hi/lo product (fp32-equivalent) or plain fp16
    int k_blocks = 1;                // K / 64 per term
    int r_in = 128, r_out = 1;       // 128-row tile = r_out outer groups x r_in inner rows
    int L_inner = 0, n_outer = 1;    // logical extents; output row = outer * L_inner + inner
    int N = 0, n_tile = 128;
    float* C = nullptr; int ldc = 0;
    int c_last_only = 0;             // store only the row inner = L_inner - 1 of every outer group: C is [n_outer][N]
    __half* Chi = nullptr; __half* Clo = nullptr; int ldh = 0;   // optional fp16 (hi, lo) copy of the output
    const float* bias = nullptr; int relu = 0;
    const float* gbias = nullptr; int gsize = 1, ldg = 0;   // per row-group bias [(row / gsize), N]
    float* stats = nullptr;          // [n_outer][ceil(L_inner / r_in)][2 column halves][2] partial (sum, sum^2) or null
    int* err = nullptr;
    // GroupNorm(1 group over an outer group's L_inner x N outputs) + shortcut + ReLU in the epilogue, output as the padded
    // fp16 (hi, lo) operand [n_outer][(L + 2) * C] of the next conv (set gn_out_hi; needs r_in >= L_inner: whole groups per tile)
    const float* gn_gamma = nullptr; const float* gn_beta = nullptr; int gn_C = 0; float gn_inv_n = 0.f;
    const __half* gn_res_hi = nullptr; const __half* gn_res_lo = nullptr;
    __half* gn_out_hi = nullptr; __half* gn_out_lo = nullptr;
    int64_t gn_ld_group = 0;
    const float* gn_res_raw = nullptr; const float* gn_res_stats = nullptr;      // second GroupNorm'd input: raw output + `stats` of another GEMM
    const float* gn_res_gamma = nullptr; const float* gn_res_beta = nullptr;
    const float* gn_up_prev = nullptr;     // + linear x2 upsampling of the coarser FPN level [n_outer][L/2][N] (fp32)
    float* gn_out_f32 = nullptr;           // fp32 output [n_outer][L][N] instead of the (hi, lo) operand
};

const char* tcg_encode_a(void* map, const __half* base, int64_t k_extent, int64_t inner, int64_t outer,
                         int64_t inner_stride_elems, int64_t outer_stride_elems, int r_in, int r_out);
const char* tcg_encode_w(void* map, const __half* base, int64_t k_total, int64_t n_rows, int n_tile);
const char* tcg_launch(const TcGemm& p, int sm_count, cudaStream_t st);

struct TcApply {
    const float* raw = nullptr; const float* stats = nullptr; const float* gamma = nullptr; const float* beta = nullptr;
    const float* res_raw = nullptr; const float* res_stats = nullptr; const float* res_gamma = nullptr; const float* res_beta = nullptr;
    const __half* res_hi = nullptr; const __half* res_lo = nullptr;
    __half* out_hi = nullptr; __half* out_lo = nullptr; float* out_f32 = nullptr;
    const float* up_prev = nullptr;  // FPN top-down step: + linear x2 upsampling of the coarser level [A][L/2][C] (fp32)
    int last_only = 0;               // raw is [A][C] = time step L-1 only (statistics still over L*C); out_f32 is [A][C]
    int A = 0, L = 0, C = 0, relu = 0;
};
void tcg_actor_prep(const float* actors, __half* hi, __half* lo, int A, cudaStream_t st);
void tcg_gn_apply(const TcApply& q, cudaStream_t st);

// ---- ActorNet (reference network.py:12-61) on the GEMM engine --------------------------------
struct ActorTcConv {
    __half* W = nullptr;     // [fold*Cout][2*Kpad] fp16 = [hi | lo], k index = tap * Cin_pad + ci
    int Cout = 0, Cin_pad = 0, ksize = 3, Kpad = 0;
    int stride = 1;          // time stride of the convolution
    int fold = 1;            // `fold` consecutive output steps form one GEMM row: N = fold*Cout, taps = (fold-1)*stride + ksize
    alignas(64) unsigned char wmap[128];
};
struct ActorTc {
    std::map<std::string, ActorTcConv> conv;     // keyed by the reference weight name
    std::map<std::string, const float*> vec;     // GN affine parameters (device fp32, owned by the ctx arena)
    int* d_err = nullptr;
    bool ready = false;
    bool gn_fused = true;      // GroupNorm of bn1 / identity-shortcut bn2 in the conv GEMM's epilogue (false: separate apply pass)
};
// host fp32 weights by reference key -> packed device copies
const char* actor_tc_pack(ActorTc& a, const std::map<std::string, std::vector<float>>& host,
                          const std::map<std::string, const float*>& dev);
void actor_tc_free(ActorTc& a);
int actor_tc_fold_weights(const float* w, int Cout, int Cin, int Cin_pad, int ks, int stride, int fold, std::vector<float>& out);   // -> Kpad
int64_t actor_tc_ws_bytes(int A);
// actors [A,14,48] fp32 -> out [A,128]
const char* actor_tc_run(ActorTc& a, const float* actors, int A, void* ws, float* out, int sm_count, cudaStream_t st);

// ---- LaneNet (reference network.py:64-121) as one persistent tcgen05 kernel (lane_tc.cu) --------------------
struct LaneTc {
    __half* W = nullptr;       // 22 matrices [128][128] fp16: per block fc1.0, fc1.3, fc2.0 (h half), fc2.0 (max half), fc2.3 as (hi, lo); proj (hi, lo)
    float* params = nullptr;   // [31][128] biases / LayerNorm parameters in the order the kernel consumes them
    int* d_err = nullptr;
    alignas(64) unsigned char wmap[128];
    bool ready = false;
};
const char* lane_tc_pack(LaneTc& l, const std::map<std::string, std::vector<float>>& host);
void lane_tc_free(LaneTc& l);
// lanes [Lp * 10, 16] fp32 node features -> out [Lp, 128]
const char* lane_tc_run(LaneTc& l, const float* lanes, int Lp, float* out, int sm_count, cudaStream_t st);

// ---- token-side tail of a rela-fusion layer + head of the next one as one persistent tcgen05 kernel (node_tc.cu) ----
struct NodeChainLayer {
    __half* W = nullptr;       // 10 or 16 matrices [128][128] fp16, (hi, lo) in the order the kernel streams them
    float* params = nullptr;   // [11][128]
    int n_mats = 0;
    alignas(64) unsigned char wmap[128];
};
struct NodeChain {
    NodeChainLayer layer[6];
    int* d_err = nullptr;
    bool ready = false;
};
const char* node_chain_pack(NodeChain& n, int l, const float* Wo, const float* bo, const float* n2g, const float* n2b, const float* W1,
                            const float* b1, const float* W2, const float* b2, const float* n3g, const float* n3b,
                            const float* Wstq_next, const float* bstq_next);
void node_chain_free(NodeChain& n);
// x [rows,128] <- LN3(x1 + FFN(x1)), x1 = LN2(x + out_proj(attn)); stq_next [rows,384] = next layer's S | T | q (or null)
const char* node_chain_run(NodeChain& n, int l, const __half* ah, const __half* al, float* x, float* stq_next, int64_t rows,
                           int sm_count, cudaStream_t st);

}  // namespace mind
