// Tensor-core (tcgen05 / TMEM / TMA) rela-fusion layer: interface used by mind_api.cu.
#pragma once
#include "kernels.h"
#include <vector>

namespace mind {

// host pointers to the fp32 parameters of one RelaFusionLayer (reference network.py:124-232)
struct TcHostLayer {
    const float *Wmem, *mem_g, *mem_b;          // proj_memory: [128,384], LN
    const float *Wpe, *bpe, *pe_g, *pe_b;       // proj_edge (null on the last layer)
    const float *ne_g, *ne_b;                   // norm_edge
    const float *Win, *bin;                     // MHA in_proj [384,128], [384]
};

// one work item: scene b, queries j0..j0+15, n tokens, key chunks [ch0, ch1) of 8 keys; slot >= 0: this item is one
// part of a key-split item and parks its un-normalised softmax state in part_buf[slot] instead of writing the attention
// output.  mode 1 = single-query item (the last query block of a scene with n % 16 == 1, e.g. the cls token of a
// 32 x 128 scene): query j0 only, chunks of 128 keys, tile row = key.
// The work list starts with one range header per CTA: {b = first item, j0 = one past the last item} (indices into the list).
struct TcWork { int32_t b, j0, n, ch0, ch1, slot, mode, pad1; };
struct TcMerge { int32_t b, j0, n, slot0, nparts, pad0, pad1, pad2; };   // merge job of one key-split item

struct TcLayerDev {
    __half* Wcat = nullptr;     // [512][128] fp16 rows: W_e | W_pe | W_k | W_v   (K-major B operands)
    float* params = nullptr;    // [8][128] fp32: mem_g, mem_b, bpe, pe_g, pe_b, ne_g, ne_b, bv
    int has_edge = 0;
    alignas(64) unsigned char wmap[128];   // CUtensorMap over Wcat
};

// per-forward state of the fused layers: descriptor tables on the device and the tensor maps built over the caller's
// buffers.  Kept separate from the weights so that a captured CUDA graph can own a private copy (mind_api.cu).
struct TcForwardState {
    TcWork* d_work = nullptr; int work_cap = 0; int n_work = 0;
    TcMerge* d_merge = nullptr; int merge_cap = 0; int n_merge = 0;
    float* d_part = nullptr; int part_cap = 0;          // [slots][16 j][144]: acc[128] | m[8] | l[8]
    int grid = 0;                          // CTAs of the fused kernel = range headers at the front of d_work
    alignas(64) unsigned char emap[128];   // CUtensorMap over the edge stream, box [64 c][16 j][8 i]
    alignas(64) unsigned char emapq[128];  // same tensor, box [64 c][1 j][128 i] (single-query items)
    const void* emap_ptr = nullptr; int emap_B = 0, emap_N = 0;
    alignas(64) unsigned char tmap[128];   // CUtensorMap over stq (fp32 [B*Nmax, 384]): T rows of a tile
    const void* tmap_ptr = nullptr; int64_t tmap_rows = 0;
    int B = 0, Nmax = 0;
};

struct TcWeights : TcForwardState {
    TcLayerDev layer[6];
    bool packed = false;
    int* d_err = nullptr;                  // device alias of h_err
    volatile int* h_err = nullptr;         // mapped host word: protocol error code of a trapped launch
    int sm_count = 148;
};
void tc_free_forward_state(TcForwardState& f);

// all return nullptr on success, else an error string
const char* tc_pack_weights(TcWeights& w, const TcHostLayer (&hl)[6]);
void tc_free(TcWeights& w);
// static schedule (host only): range headers + work items, merge jobs of the key-split items
void tc_build_schedule(const int* n_tokens, int B, int sm_count, std::vector<TcWork>& work, std::vector<TcMerge>& merges,
                       int& n_slots, int& grid);
// per forward: work list + tensor map over edge16 [B,Nmax,Nmax,128] fp16; scenes with fewer than min_tokens tokens get
// no work items (they take the exact tier, pair_x3.cu)
const char* tc_prepare(TcWeights& w, const std::vector<SceneDesc>& sd, int B, int Nmax, int min_tokens, __half* edge16, HostStage& stage, cudaStream_t st);
// One fused layer over the whole batch: updates edge16 in place (layers 0-4), reads STQ
// [B*Nmax,384] (S | T | q/4), writes the attention output (before out-proj) as an fp16 (hi, lo) pair [B*Nmax,128].
const char* tc_fusion_layer(TcWeights& w, int layer, const float* stq, __half* attn_hi, __half* attn_lo, int sm_count, cudaStream_t st);
// bring-up self test: D[128,128] = A[128,128] . W[128,128]^T through TMA + tcgen05 + TMEM
const char* tc_selftest(const float* A_host, const float* W_host, float* D_host);

}  // namespace mind
