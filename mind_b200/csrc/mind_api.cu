// C ABI of libmind_b200.so (see include/mind_b200.h).  Host-side orchestration only: weight
// packing, workspace carving, launch sequence of the scene-prediction forward pass
// (reference planners/mind/networks/network.py:582-595).
#include "../../include/mind_b200.h"
#include "kernels.h"
#include "fusion_tc.h"
#include "tc_gemm.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace mind;

static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
#define CUDA_OK(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
    } while (0)

struct FusionLayerW {
    const float *We; int ldWe;                    // [128, 128] view of proj_memory.0.weight[:, :128]
    const float *Wstq, *bstq;                     // [384,128], [384]: S | T(+b_mem) | q*0.25
    const float *mem_g, *mem_b;                   // proj_memory.1
    const float *Wpe, *bpe, *pe_g, *pe_b, *ne_g, *ne_b;   // proj_edge / norm_edge (null on last layer)
    const float *Wkv, *bkv;                       // [256,128] rows k|v of in_proj
    const float *Wo, *bo, *n2_g, *n2_b, *W1, *b1, *W2, *b2, *n3_g, *n3_b;
};

// CUDA-event ranges per stage tag, recorded on the caller's stream when option "profile" is on
struct Profiler {
    bool on = false;
    std::map<std::string, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>> ranges;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    cudaEvent_t begin(cudaStream_t st) { cudaEvent_t e = get(); cudaEventRecord(e, st); return e; }
    void end(const char* tag, cudaEvent_t b, cudaStream_t st) { cudaEvent_t e = get(); cudaEventRecord(e, st); ranges[tag].push_back({b, e}); }
};
#define PROF_BEGIN() cudaEvent_t pb_ = c->prof.on ? c->prof.begin(st) : nullptr
#define PROF_NEXT(tag) do { if (c->prof.on) { c->prof.end(tag, pb_, st); pb_ = c->prof.begin(st); } } while (0)

const char* HostStage::begin() {
    cur = (cur + 1) % kSlots;
    Slot& s = slot[cur];
    if (!s.done && cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) return "cudaEventCreate(stage) failed";
    if (s.pending) {
        if (cudaEventSynchronize(s.done) != cudaSuccess) return "cudaEventSynchronize(stage) failed";
        s.pending = false;
    }
    s.used = 0;
    return nullptr;
}
const char* HostStage::upload(void* dev_dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return nullptr;
    Slot& s = slot[cur];
    const size_t at = (s.used + 255) & ~(size_t)255;
    if (at + bytes > s.cap) {                          // grow: earlier copies out of the old block must have executed first
        const size_t ncap = std::max<size_t>((at + bytes) * 2, 1 << 20);
        char* nh = nullptr;
        if (cudaHostAlloc((void**)&nh, ncap, cudaHostAllocDefault) != cudaSuccess) return "cudaHostAlloc(stage) failed";
        if (s.host) {
            if (s.used > 0 && cudaStreamSynchronize(st) != cudaSuccess) return "cudaStreamSynchronize(stage) failed";
            cudaFreeHost(s.host);
        }
        s.host = nh; s.cap = ncap; s.used = 0;
        return upload(dev_dst, src, bytes, st);
    }
    memcpy(s.host + at, src, bytes);
    s.used = at + bytes;
    if (cudaMemcpyAsync(dev_dst, s.host + at, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return "cudaMemcpyAsync(stage) failed";
    return nullptr;
}
void HostStage::end(cudaStream_t st) {
    Slot& s = slot[cur];
    if (s.done && cudaEventRecord(s.done, st) == cudaSuccess) s.pending = true;
}
void HostStage::release() {
    for (Slot& s : slot) {
        if (s.pending) cudaEventSynchronize(s.done);
        if (s.done) cudaEventDestroy(s.done);
        if (s.host) cudaFreeHost(s.host);
        s = Slot{};
    }
}

// One captured forward (option "graph"): the whole launch sequence of a batch shape + pointer set as a CUDA graph with
// its own descriptor tables, so that a tree level at F = 1..36 scenes costs one graph launch instead of ~170 kernel
// launches (1.7 ms of host enqueue time per forward measured at F = 1).
struct GraphEntry {
    std::vector<uint64_t> key;
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
    SceneDesc* d_sd = nullptr;
    int32_t* d_actor_scene = nullptr;
    int32_t* d_small = nullptr;
    TcForwardState tcf;
    int64_t launches = 0;
    uint64_t last_use = 0;
};

struct MindCtx {
    Profiler prof;
    HostStage stage;
    std::vector<GraphEntry> graphs;
    int use_graph = 0;
    cudaStream_t cap_stream = nullptr;   // capture origin (graphs are launched into the caller's stream)
    uint64_t graph_clock = 0;
    int64_t graph_replays = 0;
    int device = 0;
    std::map<std::string, std::vector<float>> host_w;
    float* arena = nullptr;
    size_t arena_floats = 0;
    std::map<std::string, const float*> dev_w;
    bool finalized = false;
    int precision = MIND_PREC_FP32;
    int chunk_scenes = 32;
    int actor_simt = 0;           // diagnostics: force the SIMT ActorNet in f16tc mode
    ActorNetWeights an{};
    FusionLayerW fl[6]{};
    TcWeights tc{};               // fp16 packed weights / per-layer params for the tensor-core path
    ActorTc actor_tc{};           // ActorNet on the tensor-core GEMM engine
    struct LaneW { __half* W = nullptr; alignas(64) unsigned char wmap[128]; const float* bias = nullptr; };
    LaneW lane_tc[2][4];          // per aggregate block: fc1.0, fc1.3, fc2.0[:, :128], fc2.3 as [128][hi 128 | lo 128] fp16
    LaneTc lane_fused{};          // the whole LaneNet as one persistent tcgen05 kernel (lane_tc.cu)
    int lane_unfused = 0;         // diagnostics: layer-by-layer LaneNet on the GEMM engine instead
    NodeChain node_chain{};       // per layer: out-proj + LN2 + FFN + LN3 + next layer's S|T|q as one kernel (node_tc.cu)
    int node_unfused = 0;         // diagnostics: the same as 4 GEMM-engine launches + 2 LayerNorm launches
    int* lane_err = nullptr;
    struct NodeW { __half* W = nullptr; alignas(64) unsigned char wmap[128]; int N = 0, K = 0, n_tile = 128; };
    NodeW node_tc[6][4];          // per fusion layer: [S|T|q] (384x128), out-proj (128x128), linear1 (256x128), linear2 (128x256)
    // exact tier of the tensor-core mode (scenes with fewer than tc_min_tokens tokens): W_e, W_pe, W_k|W_v as 3-term operands
    NodeW pair_tc[6][3];
    // decoder: reg.0, reg.3, reg.6 (49k rows at B = 256) and actor_proj.0, actor_proj.3 as 3-term operands of the GEMM engine
    NodeW dec_tc[5];
    int decoder_simt = 0;         // diagnostics: keep the decoder's large linears on the fp32 SIMT GEMM
    float* ei_tab = nullptr;      // edge init (channel-parallel form): per-lane parameter table [32][2][7] float2
    float ei_quad[21] = {};       //   and the closed-form variance coefficients
    int tc_min_tokens = 128;
    int32_t* d_small = nullptr; int small_cap = 0;
    // descriptor tables
    SceneDesc* d_sd = nullptr; int sd_cap = 0;
    int32_t* d_actor_scene = nullptr; int as_cap = 0;
    // last forward bookkeeping for debug taps
    std::map<std::string, std::pair<const float*, int64_t>> taps;
    int sm_count = 148;
};

static void graph_entry_free(GraphEntry& e) {
    if (e.exec) cudaGraphExecDestroy(e.exec);
    if (e.d_sd) cudaFree(e.d_sd);
    if (e.d_actor_scene) cudaFree(e.d_actor_scene);
    if (e.d_small) cudaFree(e.d_small);
    tc_free_forward_state(e.tcf);
    e = GraphEntry{};
}
static void graph_cache_clear(MindCtx* c) {
    for (GraphEntry& e : c->graphs) graph_entry_free(e);
    c->graphs.clear();
}

extern "C" const char* mind_last_error(void) { return g_err; }
extern "C" const char* mind_build_info(void) { return "libmind_b200 sm_100a (fp32 SIMT + tcgen05 f16 rela-fusion)"; }

extern "C" int mind_create(MindCtx** out, int device) {
    if (!out) return fail("mind_create: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail("mind_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail("mind_create: device %d out of range (%d devices)", device, n);
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("mind_create: device sm_%d%d is not Blackwell (sm_100a required)", prop.major, prop.minor);
    MindCtx* c = new MindCtx();
    if (const char* ev = getenv("MIND_ACTOR_GN_UNFUSED")) c->actor_tc.gn_fused = atoi(ev) == 0;     // A/B switch for profiling runs
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->tc.sm_count = c->sm_count;
    *out = c;
    return 0;
}

extern "C" void mind_destroy(MindCtx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    graph_cache_clear(c);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    c->stage.release();
    if (c->arena) cudaFree(c->arena);
    if (c->d_sd) cudaFree(c->d_sd);
    if (c->d_actor_scene) cudaFree(c->d_actor_scene);
    if (c->d_small) cudaFree(c->d_small);
    tc_free(c->tc);
    actor_tc_free(c->actor_tc);
    lane_tc_free(c->lane_fused);
    node_chain_free(c->node_chain);
    for (auto& blk : c->lane_tc) for (auto& lw : blk) if (lw.W) cudaFree(lw.W);
    if (c->lane_err) cudaFree(c->lane_err);
    for (auto& lay : c->node_tc) for (auto& nw : lay) if (nw.W) cudaFree(nw.W);
    for (auto& lay : c->pair_tc) for (auto& nw : lay) if (nw.W) cudaFree(nw.W);
    for (auto& nw : c->dec_tc) if (nw.W) cudaFree(nw.W);
    if (c->ei_tab) cudaFree(c->ei_tab);
    delete c;
}

extern "C" int mind_set_weight(MindCtx* c, const char* key, const float* host, int64_t numel) {
    if (!c || !key || !host || numel <= 0) return fail("mind_set_weight: bad argument");
    c->host_w[key].assign(host, host + numel);
    c->finalized = false;
    return 0;
}

extern "C" int mind_set_option(MindCtx* c, const char* name, int64_t value) {
    if (!c || !name) return fail("mind_set_option: bad argument");
    if (!strcmp(name, "precision")) {
        if (value != MIND_PREC_FP32 && value != MIND_PREC_F16TC) return fail("precision must be 0 or 1");
        c->precision = (int)value;
    } else if (!strcmp(name, "actor_simt")) {
        c->actor_simt = value != 0;
    } else if (!strcmp(name, "actor_gn_unfused")) {
        if (c->actor_tc.gn_fused == (value != 0)) graph_cache_clear(c);
        c->actor_tc.gn_fused = value == 0;      // diagnostics: GroupNorm as a separate pass behind every conv GEMM
    } else if (!strcmp(name, "decoder_simt")) {
        if (c->decoder_simt != (value != 0)) graph_cache_clear(c);
        c->decoder_simt = value != 0;
    } else if (!strcmp(name, "node_unfused")) {
        if (c->node_unfused != (value != 0)) graph_cache_clear(c);
        c->node_unfused = value != 0;
    } else if (!strcmp(name, "lane_unfused")) {
        if (c->lane_unfused != (value != 0)) graph_cache_clear(c);
        c->lane_unfused = value != 0;
    } else if (!strcmp(name, "graph")) {
        c->use_graph = value ? 1 : 0;
        if (!c->use_graph) graph_cache_clear(c);
    } else if (!strcmp(name, "profile")) {
        c->prof.on = value != 0;
    } else if (!strcmp(name, "tc_min_tokens")) {
        // tensor-core mode: scenes with fewer tokens than this run the exact (3-term, fp32 edge) tier instead of the
        // fp16-operand fused kernel; 0 = fused kernel for every scene
        if (value < 0) return fail("tc_min_tokens must be >= 0");
        if (c->tc_min_tokens != (int)value) graph_cache_clear(c);
        c->tc_min_tokens = (int)value;
    } else if (!strcmp(name, "chunk_scenes")) {
        if (value < 1) return fail("chunk_scenes must be >= 1");
        c->chunk_scenes = (int)value;
    } else {
        return fail("unknown option %s", name);
    }
    return 0;
}

extern "C" int64_t mind_launch_count(MindCtx*) { return g_launches; }
extern "C" int64_t mind_graph_replays(MindCtx* c) { return c ? c->graph_replays : -1; }

// ------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------
namespace {
struct Builder {
    std::vector<float> flat;
    std::map<std::string, size_t> off;
    size_t add(const std::string& name, const std::vector<float>& v) {
        size_t o = (flat.size() + 63) & ~size_t(63);   // 256 B alignment
        flat.resize(o);
        flat.insert(flat.end(), v.begin(), v.end());
        off[name] = o;
        return o;
    }
};
}  // namespace

static const std::vector<float>* find(MindCtx* c, const std::string& k) {
    auto it = c->host_w.find(k);
    return it == c->host_w.end() ? nullptr : &it->second;
}

extern "C" int mind_finalize_weights(MindCtx* c) {
    if (!c) return fail("mind_finalize_weights: null ctx");
    CUDA_OK(cudaSetDevice(c->device));
    Builder b;
    for (auto& kv : c->host_w) b.add(kv.first, kv.second);

    auto need = [&](const std::string& k, size_t numel) -> const std::vector<float>* {
        const std::vector<float>* v = find(c, k);
        if (!v) { fail("missing weight '%s'", k.c_str()); return nullptr; }
        if (v->size() != numel) { fail("weight '%s' has %zu elements, expected %zu", k.c_str(), v->size(), numel); return nullptr; }
        return v;
    };
    // transposed conv filters: [co][ci][k] -> [ci][k][co]
    auto conv_t = [&](const std::string& k, int co, int ci, int ks) -> bool {
        const std::vector<float>* v = need(k, (size_t)co * ci * ks);
        if (!v) return false;
        std::vector<float> t((size_t)co * ci * ks);
        for (int o = 0; o < co; ++o)
            for (int i = 0; i < ci; ++i)
                for (int kk = 0; kk < ks; ++kk) t[((size_t)i * ks + kk) * co + o] = (*v)[((size_t)o * ci + i) * ks + kk];
        b.add(k + "#T", t);
        return true;
    };
    const int Cg[4] = {32, 64, 128, 256};
    int cin = 14;
    for (int g = 0; g < 4; ++g) {
        char p[64];
        snprintf(p, sizeof p, "actor_net.groups.%d.", g);
        std::string P(p);
        if (!conv_t(P + "0.conv1.weight", Cg[g], cin, 3)) return 1;
        if (!conv_t(P + "0.conv2.weight", Cg[g], Cg[g], 3)) return 1;
        if (!conv_t(P + "0.downsample.0.weight", Cg[g], cin, 1)) return 1;
        if (!conv_t(P + "1.conv1.weight", Cg[g], Cg[g], 3)) return 1;
        if (!conv_t(P + "1.conv2.weight", Cg[g], Cg[g], 3)) return 1;
        snprintf(p, sizeof p, "actor_net.lateral.%d.conv.weight", g);
        if (!conv_t(p, 128, Cg[g], 3)) return 1;
        cin = Cg[g];
    }
    if (!conv_t("actor_net.output.conv1.weight", 128, 128, 3)) return 1;
    if (!conv_t("actor_net.output.conv2.weight", 128, 128, 3)) return 1;

    // fused node-side projection per fusion layer: rows [Ws ; Wt ; 0.25*Wq], bias [0 ; b_mem ; 0.25*bq]
    for (int l = 0; l < 6; ++l) {
        char p[96];
        snprintf(p, sizeof p, "fusion_net.fuse_scene.fusion.%d.", l);
        std::string P(p);
        const auto* Wm = need(P + "proj_memory.0.weight", 128 * 384);
        const auto* bm = need(P + "proj_memory.0.bias", 128);
        const auto* Win = need(P + "multihead_attn.in_proj_weight", 384 * 128);
        const auto* bin = need(P + "multihead_attn.in_proj_bias", 384);
        if (!Wm || !bm || !Win || !bin) return 1;
        std::vector<float> W(384 * 128), bb(384, 0.f);
        for (int o = 0; o < 128; ++o)
            for (int k = 0; k < 128; ++k) {
                W[(size_t)o * 128 + k] = (*Wm)[(size_t)o * 384 + 128 + k];
                W[(size_t)(128 + o) * 128 + k] = (*Wm)[(size_t)o * 384 + 256 + k];
                W[(size_t)(256 + o) * 128 + k] = 0.25f * (*Win)[(size_t)o * 128 + k];
            }
        for (int o = 0; o < 128; ++o) { bb[128 + o] = (*bm)[o]; bb[256 + o] = 0.25f * (*bin)[o]; }
        b.add(P + "#Wstq", W);
        b.add(P + "#bstq", bb);
    }
    if (!need("__bezier_T", 60 * 8) || !need("__bezier_Tp", 60 * 7)) return 1;

    if (c->arena) { cudaFree(c->arena); c->arena = nullptr; }
    c->arena_floats = b.flat.size();
    CUDA_OK(cudaMalloc(&c->arena, c->arena_floats * sizeof(float)));
    CUDA_OK(cudaMemcpy(c->arena, b.flat.data(), c->arena_floats * sizeof(float), cudaMemcpyHostToDevice));
    c->dev_w.clear();
    for (auto& kv : b.off) c->dev_w[kv.first] = c->arena + kv.second;

    bool ok = true;
    auto D = [&](const std::string& k) -> const float* {
        auto it = c->dev_w.find(k);
        if (it == c->dev_w.end()) { if (ok) fail("missing weight '%s'", k.c_str()); ok = false; return nullptr; }
        return it->second;
    };
    ActorNetWeights& an = c->an;
    for (int g = 0; g < 4; ++g) {
        for (int j = 0; j < 2; ++j) {
            char p[64];
            snprintf(p, sizeof p, "actor_net.groups.%d.%d.", g, j);
            std::string P(p);
            an.g_conv1[g][j] = D(P + "conv1.weight#T");
            an.g_conv2[g][j] = D(P + "conv2.weight#T");
            an.g_bn1w[g][j] = D(P + "bn1.weight"); an.g_bn1b[g][j] = D(P + "bn1.bias");
            an.g_bn2w[g][j] = D(P + "bn2.weight"); an.g_bn2b[g][j] = D(P + "bn2.bias");
        }
        char p[64];
        snprintf(p, sizeof p, "actor_net.groups.%d.0.downsample.", g);
        std::string P(p);
        an.g_ds[g] = D(P + "0.weight#T"); an.g_dsw[g] = D(P + "1.weight"); an.g_dsb[g] = D(P + "1.bias");
        snprintf(p, sizeof p, "actor_net.lateral.%d.", g);
        P = p;
        an.lat_conv[g] = D(P + "conv.weight#T"); an.lat_w[g] = D(P + "norm.weight"); an.lat_b[g] = D(P + "norm.bias");
    }
    an.out_conv1 = D("actor_net.output.conv1.weight#T"); an.out_conv2 = D("actor_net.output.conv2.weight#T");
    an.out_bn1w = D("actor_net.output.bn1.weight"); an.out_bn1b = D("actor_net.output.bn1.bias");
    an.out_bn2w = D("actor_net.output.bn2.weight"); an.out_bn2b = D("actor_net.output.bn2.bias");

    for (int l = 0; l < 6; ++l) {
        char p[96];
        snprintf(p, sizeof p, "fusion_net.fuse_scene.fusion.%d.", l);
        std::string P(p);
        FusionLayerW& f = c->fl[l];
        f.We = D(P + "proj_memory.0.weight"); f.ldWe = 384;
        f.Wstq = D(P + "#Wstq"); f.bstq = D(P + "#bstq");
        f.mem_g = D(P + "proj_memory.1.weight"); f.mem_b = D(P + "proj_memory.1.bias");
        if (l < 5) {
            f.Wpe = D(P + "proj_edge.0.weight"); f.bpe = D(P + "proj_edge.0.bias");
            f.pe_g = D(P + "proj_edge.1.weight"); f.pe_b = D(P + "proj_edge.1.bias");
            f.ne_g = D(P + "norm_edge.weight"); f.ne_b = D(P + "norm_edge.bias");
        } else {
            f.Wpe = f.bpe = f.pe_g = f.pe_b = f.ne_g = f.ne_b = nullptr;
        }
        const float* Win = D(P + "multihead_attn.in_proj_weight");
        const float* bin = D(P + "multihead_attn.in_proj_bias");
        f.Wkv = Win ? Win + 128 * 128 : nullptr; f.bkv = bin ? bin + 128 : nullptr;
        f.Wo = D(P + "multihead_attn.out_proj.weight"); f.bo = D(P + "multihead_attn.out_proj.bias");
        f.n2_g = D(P + "norm2.weight"); f.n2_b = D(P + "norm2.bias");
        f.W1 = D(P + "linear1.weight"); f.b1 = D(P + "linear1.bias");
        f.W2 = D(P + "linear2.weight"); f.b2 = D(P + "linear2.bias");
        f.n3_g = D(P + "norm3.weight"); f.n3_b = D(P + "norm3.bias");
    }
    if (!ok) return 1;

    // tensor-core path: fp16 operand copies + per-layer parameter blocks
    {
        TcHostLayer hl[6];
        for (int l = 0; l < 6; ++l) {
            char p[96];
            snprintf(p, sizeof p, "fusion_net.fuse_scene.fusion.%d.", l);
            std::string P(p);
            hl[l].Wmem = find(c, P + "proj_memory.0.weight")->data();
            hl[l].mem_g = find(c, P + "proj_memory.1.weight")->data();
            hl[l].mem_b = find(c, P + "proj_memory.1.bias")->data();
            hl[l].Win = find(c, P + "multihead_attn.in_proj_weight")->data();
            hl[l].bin = find(c, P + "multihead_attn.in_proj_bias")->data();
            if (l < 5) {
                hl[l].Wpe = find(c, P + "proj_edge.0.weight")->data();
                hl[l].bpe = find(c, P + "proj_edge.0.bias")->data();
                hl[l].pe_g = find(c, P + "proj_edge.1.weight")->data();
                hl[l].pe_b = find(c, P + "proj_edge.1.bias")->data();
                hl[l].ne_g = find(c, P + "norm_edge.weight")->data();
                hl[l].ne_b = find(c, P + "norm_edge.bias")->data();
            } else {
                hl[l].Wpe = hl[l].bpe = hl[l].pe_g = hl[l].pe_b = hl[l].ne_g = hl[l].ne_b = nullptr;
            }
        }
        const char* err = tc_pack_weights(c->tc, hl);
        if (err) return fail("tc_pack_weights: %s", err);
        err = actor_tc_pack(c->actor_tc, c->host_w, c->dev_w);
        if (err) return fail("actor_tc_pack: %s", err);
        err = lane_tc_pack(c->lane_fused, c->host_w);
        if (err) return fail("%s", err);
        // LaneNet linears as [N=128][hi K=128 | lo K=128] fp16 operands of the GEMM engine
        const char* lnames[4] = {"fc1.0", "fc1.3", "fc2.0", "fc2.3"};
        for (int blk = 0; blk < 2; ++blk)
            for (int j = 0; j < 4; ++j) {
                const std::string P = "lane_net.aggre" + std::to_string(blk + 1) + "." + lnames[j];
                const std::vector<float>* Wv = find(c, P + ".weight");
                if (!Wv) return fail("missing %s", P.c_str());
                const int ldw = (j == 2) ? 256 : 128;
                std::vector<__half> Wp((size_t)128 * 256);
                for (int o = 0; o < 128; ++o)
                    for (int k = 0; k < 128; ++k) {
                        const float wv = (*Wv)[(size_t)o * ldw + k];
                        const __half hh = __float2half_rn(wv);
                        Wp[(size_t)o * 256 + k] = hh;
                        Wp[(size_t)o * 256 + 128 + k] = __float2half_rn(wv - __half2float(hh));
                    }
                MindCtx::LaneW& lw = c->lane_tc[blk][j];
                if (!lw.W) CUDA_OK(cudaMalloc(&lw.W, Wp.size() * sizeof(__half)));
                CUDA_OK(cudaMemcpy(lw.W, Wp.data(), Wp.size() * sizeof(__half), cudaMemcpyHostToDevice));
                if (const char* e2 = tcg_encode_w(lw.wmap, lw.W, 256, 128, 128)) return fail("lane wmap: %s", e2);
                lw.bias = (j == 2) ? nullptr : c->dev_w.at(P + ".bias");
            }
        if (!c->lane_err) { CUDA_OK(cudaMalloc(&c->lane_err, sizeof(int))); CUDA_OK(cudaMemset(c->lane_err, 0, sizeof(int))); }
        // node-side projections of every fusion layer as [N][hi K | lo K] fp16 operands of the GEMM engine
        for (int l = 0; l < 6; ++l) {
            char pfx[96];
            snprintf(pfx, sizeof pfx, "fusion_net.fuse_scene.fusion.%d.", l);
            const std::string P(pfx);
            struct Spec { const std::vector<float>* W; int N, K; float scale_from_row, scale; };
            auto pack = [&](MindCtx::NodeW& nw, const std::vector<float>& Wsrc, int N, int K, int n_tile) -> const char* {
                std::vector<__half> Wp((size_t)N * 2 * K);
                for (int o = 0; o < N; ++o)
                    for (int k = 0; k < K; ++k) {
                        const float wv = Wsrc[(size_t)o * K + k];
                        const __half hh = __float2half_rn(wv);
                        Wp[(size_t)o * 2 * K + k] = hh;
                        Wp[(size_t)o * 2 * K + K + k] = __float2half_rn(wv - __half2float(hh));
                    }
                if (!nw.W && cudaMalloc(&nw.W, Wp.size() * sizeof(__half)) != cudaSuccess) return "cudaMalloc(node W) failed";
                if (cudaMemcpy(nw.W, Wp.data(), Wp.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return "copy node W failed";
                nw.N = N; nw.K = K; nw.n_tile = n_tile;
                return tcg_encode_w(nw.wmap, nw.W, 2 * K, N, n_tile);
            };
            // fused [S | T | q/4] projection: same rows as the fp32 #Wstq pack
            const std::vector<float>* Wm = find(c, P + "proj_memory.0.weight");
            const std::vector<float>* Win = find(c, P + "multihead_attn.in_proj_weight");
            std::vector<float> Wstq((size_t)384 * 128);
            for (int o = 0; o < 128; ++o)
                for (int k = 0; k < 128; ++k) {
                    Wstq[(size_t)o * 128 + k] = (*Wm)[(size_t)o * 384 + 128 + k];
                    Wstq[(size_t)(128 + o) * 128 + k] = (*Wm)[(size_t)o * 384 + 256 + k];
                    Wstq[(size_t)(256 + o) * 128 + k] = 0.25f * (*Win)[(size_t)o * 128 + k];
                }
            const char* e3;
            if ((e3 = pack(c->node_tc[l][0], Wstq, 384, 128, 128))) return fail("node pack: %s", e3);
            if ((e3 = pack(c->node_tc[l][1], *find(c, P + "multihead_attn.out_proj.weight"), 128, 128, 128))) return fail("node pack: %s", e3);
            if ((e3 = pack(c->node_tc[l][2], *find(c, P + "linear1.weight"), 256, 128, 256))) return fail("node pack: %s", e3);
            if ((e3 = pack(c->node_tc[l][3], *find(c, P + "linear2.weight"), 128, 256, 128))) return fail("node pack: %s", e3);
            // exact tier: the N^2-row contractions W_e (proj_memory.0.weight[:, :128]), W_pe, W_k|W_v (in_proj rows 128..383)
            std::vector<float> We((size_t)128 * 128);
            for (int o = 0; o < 128; ++o)
                for (int k = 0; k < 128; ++k) We[(size_t)o * 128 + k] = (*Wm)[(size_t)o * 384 + k];
            if ((e3 = pack(c->pair_tc[l][0], We, 128, 128, 128))) return fail("pair pack: %s", e3);
            if (l < 5 && (e3 = pack(c->pair_tc[l][1], *find(c, P + "proj_edge.0.weight"), 128, 128, 128))) return fail("pair pack: %s", e3);
            std::vector<float> Wkv(Win->begin() + 128 * 128, Win->end());
            if ((e3 = pack(c->pair_tc[l][2], Wkv, 256, 128, 256))) return fail("pair pack: %s", e3);
        }
        {   // edge init: closed-form LayerNorm statistics + per-lane parameter table (simt_kernels.cu, k_edge_init_ch)
            const std::vector<float>* Wv = find(c, "fusion_net.proj_rpe_scene.0.weight");
            const std::vector<float>* bv = find(c, "fusion_net.proj_rpe_scene.0.bias");
            const std::vector<float>* gv = find(c, "fusion_net.proj_rpe_scene.1.weight");
            const std::vector<float>* ev = find(c, "fusion_net.proj_rpe_scene.1.bias");
            if (!Wv || !bv || !gv || !ev || Wv->size() != 640 || bv->size() != 128 || gv->size() != 128 || ev->size() != 128)
                return fail("missing / mis-sized fusion_net.proj_rpe_scene weights");
            std::vector<float> tab(896);
            edge_init_pack_ch(Wv->data(), bv->data(), gv->data(), ev->data(), tab.data(), c->ei_quad);
            if (!c->ei_tab) CUDA_OK(cudaMalloc(&c->ei_tab, tab.size() * sizeof(float)));
            CUDA_OK(cudaMemcpy(c->ei_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        {   // decoder linears with many rows: [N rows, padded to the column tile][hi K | lo K]
            struct DSpec { const char* key; int N, K, n_tile; };
            const DSpec ds[5] = {{"pred_scene.reg.0.weight", 128, 128, 128}, {"pred_scene.reg.3.weight", 128, 128, 128},
                                 {"pred_scene.reg.6.weight", 40, 128, 64}, {"pred_scene.actor_proj.0.weight", 384, 128, 128},
                                 {"pred_scene.actor_proj.3.weight", 768, 384, 128}};
            for (int i = 0; i < 5; ++i) {
                const std::vector<float>* Wv = find(c, ds[i].key);
                if (!Wv || Wv->size() != (size_t)ds[i].N * ds[i].K) return fail("missing / mis-sized %s", ds[i].key);
                const int K = ds[i].K, Np = ((ds[i].N + ds[i].n_tile - 1) / ds[i].n_tile) * ds[i].n_tile;
                std::vector<__half> Wp((size_t)Np * 2 * K, __float2half(0.f));
                for (int o = 0; o < ds[i].N; ++o)
                    for (int k = 0; k < K; ++k) {
                        const float wv = (*Wv)[(size_t)o * K + k];
                        const __half hh = __float2half_rn(wv);
                        Wp[(size_t)o * 2 * K + k] = hh;
                        Wp[(size_t)o * 2 * K + K + k] = __float2half_rn(wv - __half2float(hh));
                    }
                MindCtx::NodeW& nw = c->dec_tc[i];
                if (!nw.W) CUDA_OK(cudaMalloc(&nw.W, Wp.size() * sizeof(__half)));
                CUDA_OK(cudaMemcpy(nw.W, Wp.data(), Wp.size() * sizeof(__half), cudaMemcpyHostToDevice));
                nw.N = ds[i].N; nw.K = K; nw.n_tile = ds[i].n_tile;
                if (const char* e5 = tcg_encode_w(nw.wmap, nw.W, 2 * K, Np, ds[i].n_tile)) return fail("decoder wmap: %s", e5);
            }
        }
        // token-side chain kernel: layer l's out-proj / FFN / norms + layer l+1's fused [S | T | q/4] projection
        for (int l = 0; l < 6; ++l) {
            char pfx[96];
            snprintf(pfx, sizeof pfx, "fusion_net.fuse_scene.fusion.%d.", l);
            const std::string P(pfx);
            auto H = [&](const std::string& k) { return find(c, P + k)->data(); };
            const float *Wn = nullptr, *bn = nullptr;
            if (l < 5) {
                snprintf(pfx, sizeof pfx, "fusion_net.fuse_scene.fusion.%d.", l + 1);
                Wn = b.flat.data() + b.off.at(std::string(pfx) + "#Wstq");
                bn = b.flat.data() + b.off.at(std::string(pfx) + "#bstq");
            }
            if (const char* e4 = node_chain_pack(c->node_chain, l, H("multihead_attn.out_proj.weight"), H("multihead_attn.out_proj.bias"),
                                                 H("norm2.weight"), H("norm2.bias"), H("linear1.weight"), H("linear1.bias"), H("linear2.weight"),
                                                 H("linear2.bias"), H("norm3.weight"), H("norm3.bias"), Wn, bn))
                return fail("%s", e4);
        }
    }
    c->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------
// workspace carving (identical code path for sizing and for the real run)
// ------------------------------------------------------------------------------------------
namespace {
struct Carver {
    char* base;
    int64_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    template <typename T>
    T* take(int64_t n) {
        off = (off + 255) & ~int64_t(255);
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * (int64_t)sizeof(T);
        return p;
    }
};

struct Ws {
    // encoders
    float *actor_feat, *lane_in, *lx, *la, *ly, *lm, *lgb, *lane_feat;
    float *actor_p, *lane_p;
    // tokens
    float *x, *stq, *attn, *xo, *ffn;
    // exact path chunk buffers
    float *edge, *tmp, *memory, *kv;
    // tensor-core path
    __half* edge16;
    char* actor_ws;
    __half *lh[3], *ll[3];     // lane-net fp16 hi/lo operand buffers [R,128]
    float* lt32;
    __half *xh, *xl, *ah, *al, *fh, *fl;   // token state / attention output / FFN hidden as fp16 hi/lo operands
    __half *dh, *dl;                       // decoder: operand of the current large linear [A*6,128] or [A,384]
    // exact tier (compact pair grid of the small scenes): fp32 edge + operands of the 3-term GEMMs
    float *xe32, *xtmp, *xkv;
    __half *xeh, *xel, *xmh, *xml;
    // decoder
    float *actors_f, *cls_tok, *tr, *tg1, *tgt, *c1, *ce, *qkv, *att, *co, *f1, *f2, *a1, *ae, *embed, *h1, *h2, *param;
    float *k1, *k2, *logit;
};

int chunk_for(const MindCtx* c, int B) { return std::max(1, std::min(B, c->chunk_scenes)); }

// n_small scenes (each on an Ns x Ns compact pair grid) take the exact tier of the tensor-core mode, n_big the fused kernel
int64_t carve(const MindCtx* c, void* base, int B, int A, int L, int Nmax, int n_small, int Ns, int n_big, Ws& w) {
    Carver cv(base);
    const int64_t Lp = (int64_t)L + B, R = Lp * 10, TOK = (int64_t)B * Nmax;
    w.actor_feat = cv.take<float>((int64_t)A * 128);
    w.lane_in = cv.take<float>(R * 16);
    w.lx = cv.take<float>(R * 128);
    w.la = cv.take<float>(R * 128);
    w.ly = cv.take<float>(R * 128);
    w.lm = cv.take<float>(Lp * 128);
    w.lgb = cv.take<float>(Lp * 128);
    w.lane_feat = cv.take<float>(Lp * 128);
    w.actor_p = cv.take<float>((int64_t)A * 128);
    w.lane_p = cv.take<float>((int64_t)std::max(L, 1) * 128);
    w.x = cv.take<float>(TOK * 128);
    w.stq = cv.take<float>(TOK * 384);
    w.attn = cv.take<float>(TOK * 128);
    w.xo = cv.take<float>(TOK * 128);
    w.ffn = cv.take<float>(TOK * 256);
    const int64_t pairs = (int64_t)Nmax * Nmax;
    if (c->precision == MIND_PREC_FP32) {
        const int64_t C = chunk_for(c, B);
        w.edge = cv.take<float>(C * pairs * 128);
        w.tmp = cv.take<float>(C * pairs * 128);
        w.memory = cv.take<float>(C * pairs * 128);
        w.kv = cv.take<float>(C * pairs * 256);
        w.edge16 = nullptr;
        w.actor_ws = nullptr;
        for (int i = 0; i < 3; ++i) w.lh[i] = w.ll[i] = nullptr;
        w.lt32 = nullptr;
        w.xh = w.xl = w.ah = w.al = w.fh = w.fl = nullptr;
        w.dh = w.dl = nullptr;
    } else {
        w.edge = w.tmp = w.memory = w.kv = nullptr;
        w.edge16 = n_big > 0 ? cv.take<__half>((int64_t)B * pairs * 128) : nullptr;
        const int64_t xr = (int64_t)n_small * Ns * Ns;
        w.xe32 = w.xtmp = w.xkv = nullptr; w.xeh = w.xel = w.xmh = w.xml = nullptr;
        if (xr > 0) {
            w.xe32 = cv.take<float>(xr * 128); w.xtmp = cv.take<float>(xr * 128); w.xkv = cv.take<float>(xr * 256);
            w.xeh = cv.take<__half>(xr * 128 + 512); w.xel = cv.take<__half>(xr * 128 + 512);
            w.xmh = cv.take<__half>(xr * 128 + 512); w.xml = cv.take<__half>(xr * 128 + 512);
        }
        w.actor_ws = cv.take<char>(actor_tc_ws_bytes(A));
        for (int i = 0; i < 3; ++i) { w.lh[i] = cv.take<__half>(R * 128 + 512); w.ll[i] = cv.take<__half>(R * 128 + 512); }
        w.lt32 = cv.take<float>(R * 128);
        w.xh = cv.take<__half>(TOK * 128 + 512); w.xl = cv.take<__half>(TOK * 128 + 512);
        w.ah = cv.take<__half>(TOK * 128 + 512); w.al = cv.take<__half>(TOK * 128 + 512);
        w.fh = cv.take<__half>(TOK * 256 + 512); w.fl = cv.take<__half>(TOK * 256 + 512);
        w.dh = cv.take<__half>((int64_t)A * 768 + 512); w.dl = cv.take<__half>((int64_t)A * 768 + 512);
    }
    w.actors_f = cv.take<float>((int64_t)A * 128);
    w.cls_tok = cv.take<float>((int64_t)B * 128);
    w.tr = cv.take<float>((int64_t)B * 128);
    w.tg1 = cv.take<float>((int64_t)B * 128);
    w.tgt = cv.take<float>((int64_t)B * 128);
    w.c1 = cv.take<float>((int64_t)B * 384);
    w.ce = cv.take<float>((int64_t)B * 768);
    w.qkv = cv.take<float>((int64_t)B * 6 * 384);
    w.att = cv.take<float>((int64_t)B * 768);
    w.co = cv.take<float>((int64_t)B * 768);
    w.f1 = cv.take<float>((int64_t)B * 6 * 1536);
    w.f2 = cv.take<float>((int64_t)B * 768);
    w.a1 = cv.take<float>((int64_t)A * 384);
    w.ae = cv.take<float>((int64_t)A * 768);
    w.embed = cv.take<float>((int64_t)A * 768);
    w.h1 = cv.take<float>((int64_t)A * 768);
    w.h2 = cv.take<float>((int64_t)A * 768);
    w.param = cv.take<float>((int64_t)A * 6 * 40);
    w.k1 = cv.take<float>((int64_t)B * 768);
    w.k2 = cv.take<float>((int64_t)B * 768);
    w.logit = cv.take<float>((int64_t)B * 6);
    return (cv.off + 255) & ~int64_t(255);
}
}  // namespace

extern "C" int64_t mind_workspace_bytes(MindCtx* c, int32_t B, int32_t A, int32_t L, int32_t Nmax) {
    if (!c || B <= 0 || A < 0 || L < 0 || Nmax <= 0) { fail("mind_workspace_bytes: bad argument"); return -1; }
    Ws w;      // uniform batch: every scene has Nmax tokens (exact for tree levels and the benchmark batch)
    const bool small = c->precision == MIND_PREC_F16TC && Nmax < c->tc_min_tokens;
    return carve(c, nullptr, B, A, L, Nmax, small ? B : 0, small ? Nmax : 0, small ? 0 : B, w);
}

// exact requirement of one batch (ragged scenes: how many take which tier follows from the offsets)
static void tier_split(const MindCtx* c, const MindBatch* bt, int& Nmax, int& n_small, int& Ns, int& n_big) {
    Nmax = n_small = Ns = n_big = 0;
    for (int b = 0; b < bt->n_scenes; ++b) {
        const int n = (bt->actor_off[b + 1] - bt->actor_off[b]) + (bt->lane_off[b + 1] - bt->lane_off[b]) + 1;
        Nmax = std::max(Nmax, n);
        if (c->precision == MIND_PREC_F16TC && n < c->tc_min_tokens) { ++n_small; Ns = std::max(Ns, n); }
        else ++n_big;
    }
}
extern "C" int64_t mind_workspace_bytes_batch(MindCtx* c, const MindBatch* bt) {
    if (!c || !bt || bt->n_scenes <= 0 || !bt->actor_off || !bt->lane_off) { fail("mind_workspace_bytes_batch: bad argument"); return -1; }
    int Nmax, n_small, Ns, n_big;
    tier_split(c, bt, Nmax, n_small, Ns, n_big);
    Ws w;
    const int B = bt->n_scenes;
    return carve(c, nullptr, B, bt->actor_off[B] - bt->actor_off[0], bt->lane_off[B] - bt->lane_off[0], Nmax, n_small, Ns, n_big, w);
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
namespace {
struct Lin {   // Linear (+ optional LayerNorm + ReLU) helper
    MindCtx* c;
    cudaStream_t st;
    const float* W(const std::string& k) const { return c->dev_w.at(k); }
    void gemm(const float* A, int lda, const float* Wp, int ldw, const float* bias, float* C, int ldc, int64_t M, int N,
              int K, int relu = 0, const float* gbias = nullptr, int gsize = 1, int ldg = 0) const {
        GemmArgs g;
        g.A = A; g.lda = lda; g.W = Wp; g.ldw = ldw; g.bias = bias; g.C = C; g.ldc = ldc;
        g.M = (int)M; g.N = N; g.K = K; g.relu = relu; g.gbias = gbias; g.gsize = gsize; g.ldg = ldg;
        launch_gemm(g, st);
    }
    // y = ReLU(LN(x W^T + b)) with keys  <p><i>.weight/.bias and <p><i+1>.weight/.bias
    void lin_ln_relu(const std::string& p, int i, const float* x, int K, float* y, int N, int64_t M) const {
        const std::string a = p + std::to_string(i), n = p + std::to_string(i + 1);
        gemm(x, K, W(a + ".weight"), K, W(a + ".bias"), y, N, M, N, K);
        launch_layernorm(y, nullptr, W(n + ".weight"), W(n + ".bias"), y, M, N, 1, st);
    }
};
}  // namespace

static void run_lane_net(MindCtx* c, const Ws& w, int64_t Lp, cudaStream_t st) {
    Lin L{c, st};
    const int64_t R = Lp * 10;
    L.lin_ln_relu("lane_net.proj.", 0, w.lane_in, 16, w.lx, 128, R);             // network.py:108-112,118
    for (int blk = 1; blk <= 2; ++blk) {                                         // PointAggregateBlock :90-99
        const std::string P = "lane_net.aggre" + std::to_string(blk) + ".";
        L.lin_ln_relu(P + "fc1.", 0, w.lx, 128, w.la, 128, R);
        L.lin_ln_relu(P + "fc1.", 3, w.la, 128, w.ly, 128, R);                   // ly = fc1(x)
        launch_group_max(w.ly, w.lm, Lp, 10, 128, st);                           // max over the 10 nodes
        // fc2.0 on cat[h, max]: W = [Wa | Wb];  Wb.max (+bias) is a per-polyline row-group bias
        const float* W20 = L.W(P + "fc2.0.weight");
        L.gemm(w.lm, 128, W20 + 128, 256, L.W(P + "fc2.0.bias"), w.lgb, 128, Lp, 128, 128);
        L.gemm(w.ly, 128, W20, 256, nullptr, w.la, 128, R, 128, 128, 0, w.lgb, 10, 128);
        launch_layernorm(w.la, nullptr, L.W(P + "fc2.1.weight"), L.W(P + "fc2.1.bias"), w.la, R, 128, 1, st);
        L.lin_ln_relu(P + "fc2.", 3, w.la, 128, w.ly, 128, R);
        if (blk == 1) {
            launch_layernorm(w.lx, w.ly, L.W(P + "norm.weight"), L.W(P + "norm.bias"), w.lx, R, 128, 0, st);
        } else {
            launch_layernorm(w.lx, w.ly, L.W(P + "norm.weight"), L.W(P + "norm.bias"), w.la, R, 128, 0, st);
            launch_group_max(w.la, w.lane_feat, Lp, 10, 128, st);
        }
    }
}

// LaneNet with its eight [R,128]x[128,128] linears on the tcgen05 GEMM engine (3-term fp16 split, fp32-equivalent);
// LayerNorm kernels emit the (hi, lo) operands.  Same dataflow as run_lane_net (network.py:64-121).
static const char* run_lane_net_tc(MindCtx* c, const Ws& w, int64_t Lp, cudaStream_t st) {
    Lin L{c, st};
    const int64_t R = Lp * 10;
    const char* err = nullptr;
    auto tc = [&](const __half* ah, const __half* al, const MindCtx::LaneW& lw, float* C, const float* gbias) {
        if (err) return;
        alignas(64) unsigned char mh[128], ml[128];
        if ((err = tcg_encode_a(mh, ah, 128, R, 1, 128, R * 128, 128, 1))) return;
        if ((err = tcg_encode_a(ml, al, 128, R, 1, 128, R * 128, 128, 1))) return;
        TcGemm g;
        g.amap_hi = mh; g.amap_lo = ml; g.wmap = lw.wmap; g.split = 1; g.k_blocks = 2;
        g.r_in = 128; g.r_out = 1; g.L_inner = (int)R; g.n_outer = 1; g.N = 128; g.n_tile = 128;
        g.C = C; g.ldc = 128; g.bias = lw.bias; g.gbias = gbias; g.gsize = 10; g.ldg = 128; g.err = c->lane_err;
        err = tcg_launch(g, c->sm_count, st);
    };
    // proj (K = 16: SIMT), x -> fp32 + hi/lo
    L.gemm(w.lane_in, 16, L.W("lane_net.proj.0.weight"), 16, L.W("lane_net.proj.0.bias"), w.lt32, 128, R, 128, 16);
    launch_layernorm_hl(w.lt32, nullptr, L.W("lane_net.proj.1.weight"), L.W("lane_net.proj.1.bias"), w.lx, w.lh[0], w.ll[0], R, 1, st);
    for (int blk = 0; blk < 2 && !err; ++blk) {
        const std::string P = "lane_net.aggre" + std::to_string(blk + 1) + ".";
        tc(w.lh[0], w.ll[0], c->lane_tc[blk][0], w.lt32, nullptr);
        launch_layernorm_hl(w.lt32, nullptr, L.W(P + "fc1.1.weight"), L.W(P + "fc1.1.bias"), nullptr, w.lh[1], w.ll[1], R, 1, st);
        tc(w.lh[1], w.ll[1], c->lane_tc[blk][1], w.lt32, nullptr);
        launch_layernorm_hl(w.lt32, nullptr, L.W(P + "fc1.4.weight"), L.W(P + "fc1.4.bias"), w.ly, w.lh[2], w.ll[2], R, 1, st);
        launch_group_max(w.ly, w.lm, Lp, 10, 128, st);
        const float* W20 = L.W(P + "fc2.0.weight");
        L.gemm(w.lm, 128, W20 + 128, 256, L.W(P + "fc2.0.bias"), w.lgb, 128, Lp, 128, 128);
        tc(w.lh[2], w.ll[2], c->lane_tc[blk][2], w.lt32, w.lgb);
        launch_layernorm_hl(w.lt32, nullptr, L.W(P + "fc2.1.weight"), L.W(P + "fc2.1.bias"), nullptr, w.lh[1], w.ll[1], R, 1, st);
        tc(w.lh[1], w.ll[1], c->lane_tc[blk][3], w.lt32, nullptr);
        launch_layernorm(w.lt32, nullptr, L.W(P + "fc2.4.weight"), L.W(P + "fc2.4.bias"), w.ly, R, 128, 1, st);
        if (blk == 0) {
            launch_layernorm_hl(w.lx, w.ly, L.W(P + "norm.weight"), L.W(P + "norm.bias"), w.lx, w.lh[0], w.ll[0], R, 0, st);
        } else {
            launch_layernorm(w.lx, w.ly, L.W(P + "norm.weight"), L.W(P + "norm.bias"), w.la, R, 128, 0, st);
            launch_group_max(w.la, w.lane_feat, Lp, 10, 128, st);
        }
    }
    return err;
}

// node-side projections on the tcgen05 GEMM engine (3-term fp16 split): x (fp32 + hi/lo) is the running token state
static const char* node_tc_gemm(MindCtx* c, const MindCtx::NodeW& nw, const __half* ah, const __half* al, int64_t rows,
                                const float* bias, int relu, float* C, int ldc, __half* chi, __half* clo, int ldh, cudaStream_t st) {
    alignas(64) unsigned char mh[128], ml[128];
    const char* err;
    if ((err = tcg_encode_a(mh, ah, nw.K, rows, 1, nw.K, rows * nw.K, 128, 1))) return err;
    if ((err = tcg_encode_a(ml, al, nw.K, rows, 1, nw.K, rows * nw.K, 128, 1))) return err;
    TcGemm g;
    g.amap_hi = mh; g.amap_lo = ml; g.wmap = nw.wmap; g.split = 1; g.k_blocks = nw.K / 64;
    g.r_in = 128; g.r_out = 1; g.L_inner = (int)rows; g.n_outer = 1; g.N = nw.N; g.n_tile = nw.n_tile;
    g.C = C; g.ldc = ldc; g.Chi = chi; g.Clo = clo; g.ldh = ldh; g.bias = bias; g.relu = relu; g.err = c->lane_err;
    return tcg_launch(g, c->sm_count, st);
}

static const char* run_node_pre_tc(MindCtx* c, int l, const Ws& w, int64_t rows, cudaStream_t st) {
    return node_tc_gemm(c, c->node_tc[l][0], w.xh, w.xl, rows, c->fl[l].bstq, 0, w.stq, 384, nullptr, nullptr, 0, st);
}

// out-proj, LN2, FFN, LN3 (network.py:178-179,222-232); x fp32 and its hi/lo copy are updated in place
static const char* run_node_post_tc(MindCtx* c, int l, const Ws& w, int64_t rows, cudaStream_t st) {
    const FusionLayerW& f = c->fl[l];
    const char* err;
    if ((err = node_tc_gemm(c, c->node_tc[l][1], w.ah, w.al, rows, f.bo, 0, w.xo, 128, nullptr, nullptr, 0, st))) return err;
    launch_layernorm_hl(w.x, w.xo, f.n2_g, f.n2_b, w.x, w.xh, w.xl, rows, 0, st);
    if ((err = node_tc_gemm(c, c->node_tc[l][2], w.xh, w.xl, rows, f.b1, 1, nullptr, 0, w.fh, w.fl, 256, st))) return err;
    if ((err = node_tc_gemm(c, c->node_tc[l][3], w.fh, w.fl, rows, f.b2, 0, w.xo, 128, nullptr, nullptr, 0, st))) return err;
    launch_layernorm_hl(w.x, w.xo, f.n3_g, f.n3_b, w.x, w.xh, w.xl, rows, 0, st);
    return nullptr;
}

// node-side tail of a rela-fusion layer on token rows [r0, r0+rows): out-proj, LN2, FFN, LN3
// (network.py:178-179,222-232).  x is updated in place.
static void run_node_post(MindCtx* c, const FusionLayerW& f, const Ws& w, int64_t r0, int64_t rows, cudaStream_t st) {
    Lin L{c, st};
    float* x = w.x + r0 * 128;
    float* attn = w.attn + r0 * 128;
    float* xo = w.xo + r0 * 128;
    float* ffn = w.ffn + r0 * 256;
    L.gemm(attn, 128, f.Wo, 128, f.bo, xo, 128, rows, 128, 128);
    launch_layernorm(x, xo, f.n2_g, f.n2_b, x, rows, 128, 0, st);
    L.gemm(x, 128, f.W1, 128, f.b1, ffn, 256, rows, 256, 128, 1);
    L.gemm(ffn, 256, f.W2, 256, f.b2, xo, 128, rows, 128, 256);
    launch_layernorm(x, xo, f.n3_g, f.n3_b, x, rows, 128, 0, st);
}

static const char* run_decoder(MindCtx* c, const Ws& w, const MindBatch* bt, const MindOutputs* out, int B, int A,
                        const float* tgt_feat, cudaStream_t st) {
    Lin L{c, st};
    const std::string P = "pred_scene.";
    // tgt = proj_tgt(cat[tgt_feat, proj_rpe(tgt_rpe)])   (network.py:491-495)
    L.lin_ln_relu(P + "proj_rpe.", 0, bt->tgt_rpe, 20, w.tr, 128, B);
    const float* Wt0 = L.W(P + "proj_tgt.0.weight");
    L.gemm(w.tr, 128, Wt0 + 128, 256, L.W(P + "proj_tgt.0.bias"), w.tg1, 128, B, 128, 128);
    L.gemm(tgt_feat, 128, Wt0, 256, nullptr, w.tgt, 128, B, 128, 128, 0, w.tg1, 1, 128);
    launch_layernorm(w.tgt, nullptr, L.W(P + "proj_tgt.1.weight"), L.W(P + "proj_tgt.1.bias"), w.tgt, B, 128, 1, st);
    L.lin_ln_relu(P + "proj_tgt.", 3, w.tgt, 128, w.tg1, 128, B);            // tg1 = tgt [B,128]
    // cls_embed = ctx_sat(ctx_proj(cls).view(6,1,128))   (:501-502)
    L.lin_ln_relu(P + "ctx_proj.", 0, w.cls_tok, 128, w.c1, 384, B);
    L.lin_ln_relu(P + "ctx_proj.", 3, w.c1, 384, w.ce, 768, B);              // ce rows = [B*6,128]
    for (int l = 0; l < 2; ++l) {
        const std::string Q = P + "ctx_sat.layers." + std::to_string(l) + ".";
        L.gemm(w.ce, 128, L.W(Q + "self_attn.in_proj_weight"), 128, L.W(Q + "self_attn.in_proj_bias"), w.qkv, 384,
               (int64_t)B * 6, 384, 128);
        launch_mode_attention(w.qkv, w.att, B, st);
        L.gemm(w.att, 128, L.W(Q + "self_attn.out_proj.weight"), 128, L.W(Q + "self_attn.out_proj.bias"), w.co, 128,
               (int64_t)B * 6, 128, 128);
        launch_layernorm(w.ce, w.co, L.W(Q + "norm1.weight"), L.W(Q + "norm1.bias"), w.ce, (int64_t)B * 6, 128, 0, st);
        L.gemm(w.ce, 128, L.W(Q + "linear1.weight"), 128, L.W(Q + "linear1.bias"), w.f1, 1536, (int64_t)B * 6, 1536, 128, 1);
        L.gemm(w.f1, 1536, L.W(Q + "linear2.weight"), 1536, L.W(Q + "linear2.bias"), w.f2, 128, (int64_t)B * 6, 128, 1536);
        launch_layernorm(w.ce, w.f2, L.W(Q + "norm2.weight"), L.W(Q + "norm2.bias"), w.ce, (int64_t)B * 6, 128, 0, st);
    }
    // The linears over actor rows (A) and actor x mode rows (6A) run as 3-term fp16 products on the GEMM engine in the
    // tensor-core mode (fp32-equivalent, like the token-side projections): operand split -> k_tc_gemm -> LayerNorm
    const bool dtc = c->precision == MIND_PREC_F16TC && !c->decoder_simt && w.dh;
    const char* derr = nullptr;
    auto tc_lin = [&](int wi, const float* x32, int64_t rows, const std::string& bias_key, float* y, int ldy) {
        const MindCtx::NodeW& nw = c->dec_tc[wi];
        if (x32) launch_split_hl(x32, w.dh, w.dl, rows * nw.K, st);
        if (!derr) derr = node_tc_gemm(c, nw, w.dh, w.dl, rows, L.W(bias_key), 0, y, ldy, nullptr, nullptr, 0, st);
    };
    // actor_embed = actor_proj(actors).view(Na,6,128)   (:504)
    if (dtc) {
        tc_lin(3, w.actors_f, A, P + "actor_proj.0.bias", w.a1, 384);
        launch_layernorm(w.a1, nullptr, L.W(P + "actor_proj.1.weight"), L.W(P + "actor_proj.1.bias"), w.a1, A, 384, 1, st);
        tc_lin(4, w.a1, A, P + "actor_proj.3.bias", w.ae, 768);
        launch_layernorm(w.ae, nullptr, L.W(P + "actor_proj.4.weight"), L.W(P + "actor_proj.4.bias"), w.ae, A, 768, 1, st);
    } else {
        L.lin_ln_relu(P + "actor_proj.", 0, w.actors_f, 128, w.a1, 384, A);
        L.lin_ln_relu(P + "actor_proj.", 3, w.a1, 384, w.ae, 768, A);
    }
    launch_embed_combine(w.ce, w.ae, w.tg1, c->d_actor_scene, w.embed, A, st);   // (:506-510)
    // cls head (:512,547-548)
    L.lin_ln_relu(P + "cls.", 0, w.ce, 128, w.k1, 128, (int64_t)B * 6);
    L.lin_ln_relu(P + "cls.", 3, w.k1, 128, w.k2, 128, (int64_t)B * 6);
    L.gemm(w.k2, 128, L.W(P + "cls.6.weight"), 128, L.W(P + "cls.6.bias"), w.logit, 1, (int64_t)B * 6, 1, 128);
    launch_softmax6(w.logit, out->cls, B, st);
    // reg head + Bezier (:515-523,545)
    float* param = out->param ? out->param : w.param;
    if (dtc) {     // LayerNorm + ReLU write the next product's (hi, lo) operand directly
        const int64_t R6 = (int64_t)A * 6;
        tc_lin(0, w.embed, R6, P + "reg.0.bias", w.h1, 128);
        launch_layernorm_hl(w.h1, nullptr, L.W(P + "reg.1.weight"), L.W(P + "reg.1.bias"), nullptr, w.dh, w.dl, R6, 1, st);
        tc_lin(1, nullptr, R6, P + "reg.3.bias", w.h2, 128);
        launch_layernorm_hl(w.h2, nullptr, L.W(P + "reg.4.weight"), L.W(P + "reg.4.bias"), nullptr, w.dh, w.dl, R6, 1, st);
        tc_lin(2, nullptr, R6, P + "reg.6.bias", param, 40);
    } else {
        L.lin_ln_relu(P + "reg.", 0, w.embed, 128, w.h1, 128, (int64_t)A * 6);
        L.lin_ln_relu(P + "reg.", 3, w.h1, 128, w.h2, 128, (int64_t)A * 6);
        L.gemm(w.h2, 128, L.W(P + "reg.6.weight"), 128, L.W(P + "reg.6.bias"), param, 40, (int64_t)A * 6, 40, 128);
    }
    launch_bezier(param, L.W("__bezier_T"), L.W("__bezier_Tp"), out->reg, out->vel, out->cov_vel, A * 6, st);
    return derr;
}

extern "C" int mind_forward(MindCtx* c, const MindBatch* bt, const MindOutputs* out, void* workspace,
                            int64_t workspace_bytes, void* cuda_stream) {
    if (!c || !bt || !out || !workspace) return fail("mind_forward: null argument");
    if (!c->finalized) return fail("mind_forward: weights not finalized (call mind_finalize_weights)");
    const int B = bt->n_scenes;
    if (B <= 0 || !bt->actor_off || !bt->lane_off) return fail("mind_forward: empty batch or missing offsets");
    if (!bt->rpe && !(bt->ctrs && bt->vecs)) return fail("mind_forward: need rpe pointers or ctrs+vecs");
    if (!out->cls || !out->reg || !out->vel) return fail("mind_forward: cls/reg/vel outputs are required");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CUDA_OK(cudaSetDevice(c->device));
    const int A = bt->actor_off[B] - bt->actor_off[0];
    const int Ltot = bt->lane_off[B] - bt->lane_off[0];
    if (bt->actor_off[0] != 0 || bt->lane_off[0] != 0) return fail("mind_forward: offsets must start at 0");
    int Nmax = 0;
    std::vector<SceneDesc> sd(B);
    std::vector<int32_t> actor_scene((size_t)std::max(A, 1));
    std::vector<int32_t> small_ids;      // tensor-core mode: scenes that take the exact tier
    int Ns = 0;
    for (int b = 0; b < B; ++b) {
        const int na = bt->actor_off[b + 1] - bt->actor_off[b], nl = bt->lane_off[b + 1] - bt->lane_off[b];
        if (na <= 0 || nl < 0) return fail("mind_forward: scene %d has %d actors / %d lanes", b, na, nl);
        sd[b].actor_off = bt->actor_off[b]; sd[b].lane_off = bt->lane_off[b];
        sd[b].n_actor = na; sd[b].n_lane = nl;
        sd[b].geom_off = bt->actor_off[b] + bt->lane_off[b];
        sd[b].pad_ = 0;
        sd[b].rpe = bt->rpe ? bt->rpe[b] : nullptr;
        if (bt->rpe && !bt->rpe[b]) return fail("mind_forward: rpe[%d] is null", b);
        Nmax = std::max(Nmax, na + nl + 1);
        if (c->precision == MIND_PREC_F16TC && na + nl + 1 < c->tc_min_tokens) { small_ids.push_back(b); Ns = std::max(Ns, na + nl + 1); }
        for (int a = 0; a < na; ++a) actor_scene[(size_t)bt->actor_off[b] + a] = b;
    }
    const int n_small = (int)small_ids.size(), n_big = B - n_small;
    Ws w;
    const int64_t needb = carve(c, workspace, B, A, Ltot, Nmax, n_small, Ns, n_big, w);
    if (needb > workspace_bytes) return fail("mind_forward: workspace %lld B < required %lld B", (long long)workspace_bytes, (long long)needb);
    if ((((uintptr_t)workspace) & 255) != 0) return fail("mind_forward: workspace must be 256-byte aligned");

    // ---- CUDA-graph path (option "graph", tensor-core precision): exact key = shape + every pointer involved ----
    const bool can_graph = c->use_graph && c->precision == MIND_PREC_F16TC && !c->actor_simt && !c->prof.on;
    GraphEntry* ge = nullptr;
    bool capture = false;
    if (can_graph) {
        std::vector<uint64_t> key;
        key.reserve(16 + 3 * (size_t)B);
        auto kp = [&](const void* p) { key.push_back((uint64_t)(uintptr_t)p); };
        key.push_back((uint64_t)B); key.push_back((uint64_t)A); key.push_back((uint64_t)Ltot);
        key.push_back((uint64_t)c->precision | ((uint64_t)c->tc_min_tokens << 8));
        kp(bt->actors); kp(bt->lanes); kp(bt->ctrs); kp(bt->vecs); kp(bt->tgt_nodes); kp(bt->tgt_rpe);
        kp(out->cls); kp(out->reg); kp(out->vel); kp(out->cov_vel); kp(out->param); kp(workspace); kp(cuda_stream);
        for (int b = 0; b <= B; ++b) key.push_back(((uint64_t)(uint32_t)bt->actor_off[b] << 32) | (uint32_t)bt->lane_off[b]);
        if (bt->rpe) for (int b = 0; b < B; ++b) kp(bt->rpe[b]);
        for (GraphEntry& e : c->graphs) if (e.key == key) { ge = &e; break; }
        if (ge && ge->exec) {                                   // replay
            ge->last_use = ++c->graph_clock;
            CUDA_OK(cudaGraphLaunch(ge->exec, st));
            g_launches += ge->launches;
            ++c->graph_replays;
            c->taps.clear();
            c->taps["actor_feat"] = {w.actor_feat, (int64_t)A * 128};
            c->taps["lane_feat"] = {w.lane_feat, ((int64_t)Ltot + B) * 128};
            c->taps["actors_fused"] = {w.actors_f, (int64_t)A * 128};
            c->taps["cls_tok"] = {w.cls_tok, (int64_t)B * 128};
            c->taps["tokens"] = {w.x, (int64_t)B * Nmax * 128};
            return 0;
        }
        if (!ge) {                                              // first sight: plain run (lazy one-time initialisation happens here)
            if (c->graphs.size() >= 32) {
                size_t lru = 0;
                for (size_t i = 1; i < c->graphs.size(); ++i) if (c->graphs[i].last_use < c->graphs[lru].last_use) lru = i;
                graph_entry_free(c->graphs[lru]);
                c->graphs.erase(c->graphs.begin() + (long)lru);
            }
            c->graphs.emplace_back();
            c->graphs.back().key = std::move(key);
            c->graphs.back().seen = 1;
            c->graphs.back().last_use = ++c->graph_clock;
        } else {
            capture = true;                                     // second sight: capture with entry-owned descriptor tables
            ge->last_use = ++c->graph_clock;
        }
    }
    // the context's own tables are set aside while a capture builds the entry's private ones
    SceneDesc* keep_sd = c->d_sd; const int keep_sd_cap = c->sd_cap;
    int32_t* keep_as = c->d_actor_scene; const int keep_as_cap = c->as_cap;
    int32_t* keep_small = c->d_small; const int keep_small_cap = c->small_cap;
    const TcForwardState keep_tcf = c->tc;
    if (capture) {
        c->d_sd = nullptr; c->sd_cap = 0; c->d_actor_scene = nullptr; c->as_cap = 0; c->d_small = nullptr; c->small_cap = 0;
        (TcForwardState&)c->tc = TcForwardState{};
    }
    auto restore_tables = [&](bool adopt) {
        if (!capture) return;
        if (adopt) {
            ge->d_sd = c->d_sd; ge->d_actor_scene = c->d_actor_scene; ge->d_small = c->d_small; ge->tcf = c->tc;
        } else {
            if (c->d_sd) cudaFree(c->d_sd);
            if (c->d_actor_scene) cudaFree(c->d_actor_scene);
            if (c->d_small) cudaFree(c->d_small);
            tc_free_forward_state(c->tc);
        }
        c->d_sd = keep_sd; c->sd_cap = keep_sd_cap; c->d_actor_scene = keep_as; c->as_cap = keep_as_cap;
        c->d_small = keep_small; c->small_cap = keep_small_cap;
        (TcForwardState&)c->tc = keep_tcf;
    };

    // ---- descriptor tables: allocation + asynchronous upload (never part of a captured region) ----
    auto prepare = [&]() -> int {
        if (c->sd_cap < B) {
            if (c->d_sd) cudaFree(c->d_sd);
            c->d_sd = nullptr; c->sd_cap = 0;
            CUDA_OK(cudaMalloc(&c->d_sd, sizeof(SceneDesc) * (size_t)B));
            c->sd_cap = B;
        }
        if (c->as_cap < A) {
            if (c->d_actor_scene) cudaFree(c->d_actor_scene);
            c->d_actor_scene = nullptr; c->as_cap = 0;
            CUDA_OK(cudaMalloc(&c->d_actor_scene, sizeof(int32_t) * (size_t)A));
            c->as_cap = A;
        }
        if (const char* e = c->stage.begin()) return fail("mind_forward: %s", e);
        if (const char* e = c->stage.upload(c->d_sd, sd.data(), sizeof(SceneDesc) * (size_t)B, st)) return fail("mind_forward: %s", e);
        if (const char* e = c->stage.upload(c->d_actor_scene, actor_scene.data(), sizeof(int32_t) * (size_t)A, st)) return fail("mind_forward: %s", e);
        if (n_small > 0) {
            if (c->small_cap < n_small) {
                if (c->d_small) cudaFree(c->d_small);
                c->d_small = nullptr; c->small_cap = 0;
                CUDA_OK(cudaMalloc(&c->d_small, sizeof(int32_t) * (size_t)n_small));
                c->small_cap = n_small;
            }
            if (const char* e = c->stage.upload(c->d_small, small_ids.data(), sizeof(int32_t) * (size_t)n_small, st)) return fail("mind_forward: %s", e);
        }
        if (c->precision == MIND_PREC_F16TC)
            if (const char* perr = tc_prepare(c->tc, sd, B, Nmax, c->tc_min_tokens, w.edge16, c->stage, st)) return fail("tc_prepare: %s", perr);
        c->stage.end(st);
        return 0;
    };
    if (int rc = prepare()) { restore_tables(false); return rc; }

    cudaStream_t body_stream = st;       // a capture records on the context's own stream: the caller's may be the legacy one
    auto body = [&]() -> int {
    cudaStream_t st = body_stream;
    Lin L{c, st};
    PROF_BEGIN();
    // ---- encoders -------------------------------------------------------------------------
    if (c->precision == MIND_PREC_F16TC && !c->actor_simt) {                                // network.py:586
        if (const char* e = actor_tc_run(c->actor_tc, bt->actors, A, w.actor_ws, w.actor_feat, c->sm_count, st))
            return fail("actor_tc_run: %s", e);
    } else {
        launch_actor_net(bt->actors, w.actor_feat, A, c->an, st);
    }
    PROF_NEXT("actor_net");
    const int64_t Lp = (int64_t)Ltot + B;
    if (Ltot > 0)
        CUDA_OK(cudaMemcpyAsync(w.lane_in, bt->lanes, sizeof(float) * (size_t)Ltot * 160, cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemcpyAsync(w.lane_in + (int64_t)Ltot * 160, bt->tgt_nodes, sizeof(float) * (size_t)B * 160,
                            cudaMemcpyDeviceToDevice, st));
    if (c->precision == MIND_PREC_F16TC && !c->actor_simt && !c->lane_unfused) {            // :587,589
        if (const char* e = lane_tc_run(c->lane_fused, w.lane_in, (int)Lp, w.lane_feat, c->sm_count, st)) return fail("lane_tc_run: %s", e);
    } else if (c->precision == MIND_PREC_F16TC && !c->actor_simt) {
        if (const char* e = run_lane_net_tc(c, w, Lp, st)) return fail("run_lane_net_tc: %s", e);
    } else {
        run_lane_net(c, w, Lp, st);
    }
    PROF_NEXT("lane_net");
    const float* tgt_feat = w.lane_feat + (int64_t)Ltot * 128;
    // ---- fusion ---------------------------------------------------------------------------
    L.lin_ln_relu("fusion_net.proj_actor.", 0, w.actor_feat, 128, w.actor_p, 128, A);       // :313
    L.lin_ln_relu("fusion_net.proj_lane.", 0, w.lane_feat, 128, w.lane_p, 128, Ltot);       // :314
    launch_scatter_tokens(w.actor_p, w.lane_p, c->d_sd, w.x, B, Nmax, st);                  // :320-324
    const float* Wr = L.W("fusion_net.proj_rpe_scene.0.weight");
    const float* br = L.W("fusion_net.proj_rpe_scene.0.bias");
    const float* gr = L.W("fusion_net.proj_rpe_scene.1.weight");
    const float* ber = L.W("fusion_net.proj_rpe_scene.1.bias");
    const int64_t pairs = (int64_t)Nmax * Nmax;
    PROF_NEXT("tokens");
    if (c->precision == MIND_PREC_FP32) {
        const int C = chunk_for(c, B);
        for (int b0 = 0; b0 < B; b0 += C) {
            const int nb = std::min(C, B - b0);
            const int64_t r0 = (int64_t)b0 * Nmax, rows = (int64_t)nb * Nmax, prow = (int64_t)nb * pairs;
            launch_edge_init_f32(c->d_sd, bt->ctrs, bt->vecs, Wr, br, gr, ber, w.edge, b0, nb, Nmax, st);
            for (int l = 0; l < 6; ++l) {
                const FusionLayerW& f = c->fl[l];
                L.gemm(w.x + r0 * 128, 128, f.Wstq, 128, f.bstq, w.stq + r0 * 384, 384, rows, 384, 128);
                L.gemm(w.edge, 128, f.We, f.ldWe, nullptr, w.tmp, 128, prow, 128, 128);
                launch_pair_memory_epi(w.tmp, w.stq, f.mem_g, f.mem_b, w.memory, b0, nb, Nmax, st);
                if (f.Wpe) {
                    L.gemm(w.memory, 128, f.Wpe, 128, f.bpe, w.tmp, 128, prow, 128, 128);
                    launch_pair_edge_epi(w.tmp, f.pe_g, f.pe_b, f.ne_g, f.ne_b, w.edge, prow, st);
                }
                L.gemm(w.memory, 128, f.Wkv, 128, f.bkv, w.kv, 256, prow, 256, 128);
                launch_pair_attention(w.kv, w.stq, c->d_sd, w.attn, b0, nb, Nmax, st);
                run_node_post(c, f, w, r0, rows, st);
            }
        }
    } else {
        if (n_big > 0)
            launch_edge_init_ch(c->d_sd, bt->ctrs, bt->vecs, c->ei_tab, c->ei_quad, w.edge16, 0, B, Nmax, c->tc_min_tokens, st);
        const int64_t xrows = (int64_t)n_small * Ns * Ns;
        launch_x3_edge_init(c->d_sd, c->d_small, bt->ctrs, bt->vecs, Wr, br, gr, ber, w.xe32, w.xeh, w.xel, n_small, Ns, st);
        const int64_t TOKR = (int64_t)B * Nmax;
        launch_split_hl(w.x, w.xh, w.xl, TOKR * 128, st);                      // token state as fp16 hi/lo operand
        CUDA_OK(cudaMemsetAsync(w.ah, 0, sizeof(__half) * (size_t)TOKR * 128, st));   // padded token rows are never written by the fused kernel
        CUDA_OK(cudaMemsetAsync(w.al, 0, sizeof(__half) * (size_t)TOKR * 128, st));
        PROF_NEXT("edge_init");
        for (int l = 0; l < 6; ++l) {
            if (l == 0 || c->node_unfused) {       // later layers: S | T | q come out of the previous layer's chain kernel
                if (const char* e = run_node_pre_tc(c, l, w, TOKR, st)) return fail("node_pre_tc(%d): %s", l, e);
                PROF_NEXT("node_pre");
            }
            if (n_big > 0) {
                const char* err = tc_fusion_layer(c->tc, l, w.stq, w.ah, w.al, c->sm_count, st);
                if (err) return fail("tc_fusion_layer(%d): %s", l, err);
                PROF_NEXT(l < 5 ? "fusion_tc" : "fusion_tc_last");
            }
            if (n_small > 0) {      // exact tier: un-fused pair pipeline, 3-term GEMMs, fp32 edge (network.py:197-226)
                const FusionLayerW& f = c->fl[l];
                const char* e;
                if ((e = node_tc_gemm(c, c->pair_tc[l][0], w.xeh, w.xel, xrows, nullptr, 0, w.xtmp, 128, nullptr, nullptr, 0, st)))
                    return fail("x3 W_e(%d): %s", l, e);
                launch_x3_memory_epi(w.xtmp, w.stq, c->d_small, f.mem_g, f.mem_b, w.xmh, w.xml, n_small, Ns, Nmax, st);
                if (f.Wpe) {
                    if ((e = node_tc_gemm(c, c->pair_tc[l][1], w.xmh, w.xml, xrows, f.bpe, 0, w.xtmp, 128, nullptr, nullptr, 0, st)))
                        return fail("x3 W_pe(%d): %s", l, e);
                    launch_x3_edge_epi(w.xtmp, f.pe_g, f.pe_b, f.ne_g, f.ne_b, w.xe32, w.xeh, w.xel, xrows, st);
                }
                if ((e = node_tc_gemm(c, c->pair_tc[l][2], w.xmh, w.xml, xrows, f.bkv, 0, w.xkv, 256, nullptr, nullptr, 0, st)))
                    return fail("x3 W_kv(%d): %s", l, e);
                launch_x3_attention(w.xkv, w.stq, c->d_sd, c->d_small, w.ah, w.al, n_small, Ns, Nmax, st);
                PROF_NEXT("fusion_x3");
            }
            if (c->node_unfused) {
                if (const char* e = run_node_post_tc(c, l, w, TOKR, st)) return fail("node_post_tc(%d): %s", l, e);
            } else {
                if (const char* e = node_chain_run(c->node_chain, l, w.ah, w.al, w.x, l < 5 ? w.stq : nullptr, TOKR, c->sm_count, st))
                    return fail("node_chain_run(%d): %s", l, e);
            }
            PROF_NEXT("node_post");
        }
    }
    launch_gather_tokens(w.x, c->d_sd, w.actors_f, w.cls_tok, B, Nmax, st);                 // :334-336
    PROF_NEXT("fusion_other");
    // ---- decoder --------------------------------------------------------------------------
    if (const char* e = run_decoder(c, w, bt, out, B, A, tgt_feat, st)) return fail("run_decoder: %s", e);
    PROF_NEXT("decoder");
    if (c->prof.on && pb_) c->prof.pool.push_back(pb_);
    CUDA_OK(cudaGetLastError());
    return 0;
    };   // body

    const int64_t Lp = (int64_t)Ltot + B;
    if (capture) {
        if (!c->cap_stream && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            restore_tables(false);
            return fail("mind_forward: cannot create the capture stream");
        }
        body_stream = c->cap_stream;
        if (cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            restore_tables(false);
            return fail("mind_forward: cudaStreamBeginCapture failed");
        } else {
            const int64_t l0 = g_launches;
            const int rc = body();
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(c->cap_stream, &graph);
            body_stream = st;
            cudaGraphExec_t exec = nullptr;
            if (rc != 0 || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                restore_tables(false);
                ge->seen = 0;
                return rc != 0 ? rc : fail("mind_forward: CUDA graph capture failed (%s)", cudaGetErrorString(ce));
            }
            cudaGraphDestroy(graph);
            ge->exec = exec;
            ge->launches = g_launches - l0;
            restore_tables(true);
            CUDA_OK(cudaGraphLaunch(exec, st));
        }
    } else {
        if (int rc = body()) return rc;
    }
    c->taps.clear();
    c->taps["actor_feat"] = {w.actor_feat, (int64_t)A * 128};
    c->taps["lane_feat"] = {w.lane_feat, Lp * 128};
    c->taps["actors_fused"] = {w.actors_f, (int64_t)A * 128};
    c->taps["cls_tok"] = {w.cls_tok, (int64_t)B * 128};
    c->taps["tokens"] = {w.x, (int64_t)B * Nmax * 128};
    return 0;
}

extern "C" int64_t mind_debug_tap(MindCtx* c, const char* name, float* dst, int64_t capacity, void* cuda_stream) {
    if (!c || !name || !dst) { fail("mind_debug_tap: bad argument"); return -1; }
    auto it = c->taps.find(name);
    if (it == c->taps.end()) { fail("mind_debug_tap: unknown tap '%s'", name); return -1; }
    if (it->second.second > capacity) { fail("mind_debug_tap: capacity %lld < %lld", (long long)capacity, (long long)it->second.second); return -1; }
    if (cudaMemcpyAsync(dst, it->second.first, sizeof(float) * (size_t)it->second.second, cudaMemcpyDeviceToDevice,
                        (cudaStream_t)cuda_stream) != cudaSuccess) { fail("mind_debug_tap: copy failed"); return -1; }
    return it->second.second;
}

// bring-up / diagnostics -----------------------------------------------------------------------
extern "C" int mind_tc_selftest(const float* A_host, const float* W_host, float* D_host) {
    const char* e = tc_selftest(A_host, W_host, D_host);
    if (e) return fail("tc_selftest: %s", e);
    return 0;
}

extern "C" int mind_debug_edge_init_pack(const float* W, const float* b, const float* gamma, const float* beta, float* tab896,
                                         float* quad21) {
    if (!W || !b || !gamma || !beta || !tab896 || !quad21) return fail("mind_debug_edge_init_pack: null argument");
    edge_init_pack_ch(W, b, gamma, beta, tab896, quad21);
    return 0;
}

extern "C" int mind_debug_conv_fold_pack(const float* w, int32_t Cout, int32_t Cin, int32_t Cin_pad, int32_t ksize, int32_t stride,
                                         int32_t fold, float* out, int64_t capacity) {
    if (!w || !out || Cout <= 0 || Cin <= 0 || Cin_pad < Cin || (ksize != 1 && ksize != 3) || stride <= 0 || fold <= 0) {
        fail("mind_debug_conv_fold_pack: bad argument");
        return -1;
    }
    std::vector<float> Wf;
    const int Kpad = actor_tc_fold_weights(w, Cout, Cin, Cin_pad, ksize, stride, fold, Wf);
    if ((int64_t)Wf.size() > capacity) { fail("mind_debug_conv_fold_pack: capacity %lld < %zu", (long long)capacity, Wf.size()); return -1; }
    memcpy(out, Wf.data(), Wf.size() * sizeof(float));
    return Kpad;
}

extern "C" int mind_debug_fusion_schedule(const int32_t* n_tokens, int32_t B, int32_t sm_count, int32_t* work_out,
                                          int32_t capacity, int32_t* info) {
    if (!n_tokens || B <= 0 || sm_count <= 0 || !info) return fail("mind_debug_fusion_schedule: bad argument");
    std::vector<TcWork> work;
    std::vector<TcMerge> merges;
    int n_slots = 0, grid = 1;
    tc_build_schedule(n_tokens, B, sm_count, work, merges, n_slots, grid);
    info[0] = (int32_t)work.size(); info[1] = grid; info[2] = (int32_t)merges.size(); info[3] = n_slots;
    if (!work_out || capacity < (int32_t)work.size()) return -1;
    static_assert(sizeof(TcWork) == 32, "TcWork is 8 int32");
    memcpy(work_out, work.data(), work.size() * sizeof(TcWork));
    return 0;
}

// ---- batched host->device staging (reference: gpu(), planners/mind/utils.py:9-20) ----
extern "C" int64_t mind_upload_packed_bytes(const int64_t* bytes, int32_t n) {
    if (!bytes || n < 0) return -1;
    int64_t off = 0;
    for (int i = 0; i < n; ++i) {
        if (bytes[i] < 0) return -1;
        off += (bytes[i] + 255) & ~(int64_t)255;
    }
    return off;
}
extern "C" int mind_upload_packed(const void* const* host_ptrs, const int64_t* bytes, int32_t n, void* dev_dst,
                                  int64_t dst_capacity, int64_t* offsets_out, void* cuda_stream) {
    if (n < 0 || (n > 0 && (!host_ptrs || !bytes || !dev_dst || !offsets_out))) return fail("mind_upload_packed: bad argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int64_t off = 0;
    for (int i = 0; i < n; ++i) {
        if (bytes[i] < 0 || (bytes[i] > 0 && !host_ptrs[i])) return fail("mind_upload_packed: bad entry %d", i);
        if (off + bytes[i] > dst_capacity) return fail("mind_upload_packed: destination too small (%lld needed at entry %d)",
                                                       (long long)(off + bytes[i]), i);
        offsets_out[i] = off;
        if (bytes[i] > 0)
            CUDA_OK(cudaMemcpyAsync((char*)dev_dst + off, host_ptrs[i], (size_t)bytes[i], cudaMemcpyHostToDevice, st));
        off += (bytes[i] + 255) & ~(int64_t)255;
    }
    return 0;
}

// synchronise the device and report a kernel-side protocol error code (0 = none)
extern "C" int mind_sync_check(MindCtx* c) {
    if (!c) return fail("mind_sync_check: null ctx");
    cudaError_t e = cudaDeviceSynchronize();
    int code = 0;
    if (c->tc.h_err) code = *c->tc.h_err;
    if (e != cudaSuccess) return fail("device error: %s (kernel code %d)", cudaGetErrorString(e), code);
    if (code) return fail("kernel reported protocol error code %d", code);
    return 0;
}

// Drain the CUDA-event ranges recorded since the last call (option "profile" = 1).  Writes a
// text table "tag total_ms count\n" into buf; synchronises the device.
extern "C" int mind_profile_read(MindCtx* c, char* buf, int64_t cap) {
    if (!c || !buf || cap <= 0) return fail("mind_profile_read: bad argument");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail("mind_profile_read: %s", cudaGetErrorString(e));
    std::string out;
    for (auto& kv : c->prof.ranges) {
        double tot = 0;
        for (auto& pr : kv.second) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pr.first, pr.second);
            tot += ms;
            c->prof.pool.push_back(pr.first);
            c->prof.pool.push_back(pr.second);
        }
        char line[160];
        snprintf(line, sizeof line, "%s %.6f %zu\n", kv.first.c_str(), tot, kv.second.size());
        out += line;
    }
    c->prof.ranges.clear();
    if ((int64_t)out.size() + 1 > cap) return fail("mind_profile_read: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}
