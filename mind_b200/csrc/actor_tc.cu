// ActorNet (reference planners/mind/networks/network.py:12-61, layers.py:36-60,140-188) as a chain
// of tensor-core GEMMs.  Activations are channel-last, zero-padded in time, stored as an fp16
// (hi, lo) pair; a k=3 conv reads its [3*C] im2col rows as overlapping TMA windows of that buffer.
// GroupNorm(1 group): the GEMM epilogue emits per-actor partial sums, the apply kernel normalises,
// adds the shortcut, applies ReLU and re-splits to fp16 hi/lo for the next conv.
#include "tc_gemm.h"
#include <cuda_fp16.h>
#include <cstdio>

namespace mind {

namespace {
struct Carve {
    char* base; int64_t off = 0;
    template <typename T> T* take(int64_t n) {
        off = (off + 255) & ~int64_t(255);
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * (int64_t)sizeof(T);
        return p;
    }
};
struct HL { __half* hi; __half* lo; };
struct Bufs {
    HL x0, t0, t1, o[4], p;              // padded channel-last activations
    __half* slack[20]; int n_slack;      // tails that overlapping windows may read (kept finite: zeroed)
    float *raw1, *raw2, *raw3, *st1, *st2, *st3, *pyr0, *pyr1;
};
int64_t carve(void* ws, int A, Bufs& b) {
    Carve c{(char*)ws};
    const int64_t slack = 1024;   // windows of the last rows may read past the end (zero weights)
    b.n_slack = 0;
    auto hl = [&](int64_t per) {
        HL h;
        h.hi = c.take<__half>(A * per + slack); h.lo = c.take<__half>(A * per + slack);
        if (ws) { b.slack[b.n_slack++] = h.hi + A * per; b.slack[b.n_slack++] = h.lo + A * per; }
        return h;
    };
    b.x0 = hl(50 * 16);
    b.t0 = hl(50 * 128); b.t1 = hl(50 * 128);
    b.o[0] = hl(50 * 32); b.o[1] = hl(26 * 64); b.o[2] = hl(14 * 128); b.o[3] = hl(8 * 256);
    b.p = hl(50 * 128);
    b.raw1 = c.take<float>((int64_t)A * 6144); b.raw2 = c.take<float>((int64_t)A * 6144); b.raw3 = c.take<float>((int64_t)A * 6144);
    b.st1 = c.take<float>((int64_t)A * 12); b.st2 = c.take<float>((int64_t)A * 12); b.st3 = c.take<float>((int64_t)A * 12);
    b.pyr0 = c.take<float>((int64_t)A * 3072); b.pyr1 = c.take<float>((int64_t)A * 3072);   // FPN levels L = 6 / 24 and L = 12
    return (c.off + 255) & ~int64_t(255);
}
char g_aerr[256];

}  // namespace

int64_t actor_tc_ws_bytes(int A) { Bufs b; return carve(nullptr, std::max(A, 1), b); }

// Conv weight [Cout][Cin][ks] -> GEMM operand [fold*Cout][Kpad] (fp32; Kpad = taps*Cin_pad rounded up to 64, taps =
// (fold-1)*stride + ks): row u*Cout + o produces output step fold*p + u of GEMM row p from the window of `taps` padded input
// rows that starts at row fold*p*stride, so its weights sit at taps u*stride + k, zero elsewhere.  fold = 1 is the plain
// im2col layout.  Host only (also behind mind_debug_conv_fold_pack for the CPU test).
int actor_tc_fold_weights(const float* w, int Cout, int Cin, int Cin_pad, int ks, int stride, int fold, std::vector<float>& out) {
    const int taps = (fold - 1) * stride + ks;
    const int Kpad = ((taps * Cin_pad + 63) / 64) * 64;
    out.assign((size_t)fold * Cout * Kpad, 0.f);
    for (int u = 0; u < fold; ++u)
        for (int o = 0; o < Cout; ++o)
            for (int i = 0; i < Cin; ++i)
                for (int k = 0; k < ks; ++k)
                    out[((size_t)u * Cout + o) * Kpad + (size_t)(u * stride + k) * Cin_pad + i] = w[((size_t)o * Cin + i) * ks + k];
    return Kpad;
}

void actor_tc_free(ActorTc& a) {
    for (auto& kv : a.conv) if (kv.second.W) cudaFree(kv.second.W);
    a.conv.clear();
    if (a.d_err) cudaFree(a.d_err);
    a.d_err = nullptr; a.ready = false;
}

const char* actor_tc_pack(ActorTc& a, const std::map<std::string, std::vector<float>>& host,
                          const std::map<std::string, const float*>& dev) {
    actor_tc_free(a);
    // Narrow layers (C = 32) fold `fold` consecutive output steps into one GEMM row: the row's window is (fold-1)*stride + ks
    // taps of the padded input, its N = fold*Cout columns are the outputs of step fold*p + u at [u*Cout, (u+1)*Cout) -- the
    // same memory as [step][Cout] -- and the weight of (u, o) sits at taps u*stride + k (zero elsewhere).  Same kernel, but
    // 128-row tiles of N = 64 / 128 instead of 32: half / a quarter of the tiles and of the operand traffic per output.
    auto add = [&](const std::string& key, int Cout, int Cin, int Cin_pad, int ks, int stride, int fold) -> const char* {
        auto it = host.find(key);
        if (it == host.end() || it->second.size() != (size_t)Cout * Cin * ks) return "actor_tc_pack: missing / mis-sized conv weight";
        ActorTcConv cv;
        cv.Cout = Cout; cv.Cin_pad = Cin_pad; cv.ksize = ks; cv.stride = stride; cv.fold = fold;
        std::vector<float> Wf;
        cv.Kpad = actor_tc_fold_weights(it->second.data(), Cout, Cin, Cin_pad, ks, stride, fold, Wf);
        const int N = fold * Cout;
        std::vector<__half> W((size_t)N * 2 * cv.Kpad);
        for (int n = 0; n < N; ++n)
            for (int kk = 0; kk < cv.Kpad; ++kk) {
                const float w = Wf[(size_t)n * cv.Kpad + kk];
                const __half h = __float2half_rn(w);
                W[(size_t)n * 2 * cv.Kpad + kk] = h;
                W[(size_t)n * 2 * cv.Kpad + cv.Kpad + kk] = __float2half_rn(w - __half2float(h));
            }
        if (cudaMalloc(&cv.W, W.size() * sizeof(__half)) != cudaSuccess) return "actor_tc_pack: cudaMalloc failed";
        cudaMemcpy(cv.W, W.data(), W.size() * sizeof(__half), cudaMemcpyHostToDevice);
        if (const char* e = tcg_encode_w(cv.wmap, cv.W, 2 * cv.Kpad, N, N)) return e;
        a.conv[key] = cv;
        return nullptr;
    };
    int fold0 = 2;
    if (const char* ev = getenv("MIND_ACTOR_FOLD")) fold0 = atoi(ev);
    if (fold0 != 1 && fold0 != 2 && fold0 != 4) return "actor_tc_pack: MIND_ACTOR_FOLD must be 1, 2 or 4";
    const int Cg[4] = {32, 64, 128, 256};
    int cin = 14, cinp = 16;
    for (int g = 0; g < 4; ++g) {
        char p[64];
        snprintf(p, sizeof p, "actor_net.groups.%d.", g);
        std::string P(p);
        const char* e;
        const int st = g == 0 ? 1 : 2, f = g == 0 ? fold0 : 1;
        if ((e = add(P + "0.conv1.weight", Cg[g], cin, cinp, 3, st, f))) return e;
        if ((e = add(P + "0.conv2.weight", Cg[g], Cg[g], Cg[g], 3, 1, f))) return e;
        if ((e = add(P + "0.downsample.0.weight", Cg[g], cin, cinp, 1, st, f))) return e;
        if ((e = add(P + "1.conv1.weight", Cg[g], Cg[g], Cg[g], 3, 1, f))) return e;
        if ((e = add(P + "1.conv2.weight", Cg[g], Cg[g], Cg[g], 3, 1, f))) return e;
        snprintf(p, sizeof p, "actor_net.lateral.%d.conv.weight", g);
        if ((e = add(p, 128, Cg[g], Cg[g], 3, 1, 1))) return e;
        cin = cinp = Cg[g];
    }
    if (const char* e = add("actor_net.output.conv1.weight", 128, 128, 128, 3, 1, 1)) return e;
    if (const char* e = add("actor_net.output.conv2.weight", 128, 128, 128, 3, 1, 1)) return e;
    for (auto& kv : dev)
        if (kv.first.rfind("actor_net.", 0) == 0 && kv.first.find('#') == std::string::npos) a.vec[kv.first] = kv.second;
    if (cudaMalloc(&a.d_err, sizeof(int)) != cudaSuccess) return "actor_tc_pack: cudaMalloc(err) failed";
    cudaMemset(a.d_err, 0, sizeof(int));
    a.ready = true;
    return nullptr;
}

const char* actor_tc_run(ActorTc& a, const float* actors, int A, void* ws, float* out, int sm_count, cudaStream_t st) {
    if (!a.ready) return "actor_tc_run: weights not packed";
    if (A <= 0) return nullptr;
    Bufs b;
    carve(ws, A, b);
    const char* err = nullptr;
    auto V = [&](const std::string& k) -> const float* {
        auto it = a.vec.find(k);
        if (it == a.vec.end()) { snprintf(g_aerr, sizeof g_aerr, "actor_tc_run: missing %s", k.c_str()); err = g_aerr; return nullptr; }
        return it->second;
    };
    // conv over a padded channel-last input [A][Lin+2][Cin_pad]: output raw [A*Lout][Cout] + stats
    // gn != nullptr: GroupNorm (+ identity shortcut) (+ ReLU) in the GEMM epilogue, output = padded (hi, lo) operand `gn_out`
    struct Gn {
        std::string key; int relu; HL out; const HL* res;
        const float* res_raw = nullptr; const float* res_stats = nullptr; std::string res_key;   // second GroupNorm'd input (conv shortcut)
        const float* up_prev = nullptr; float* out_f32 = nullptr;                                // FPN step / fp32 output
    };
    auto conv = [&](const std::string& key, HL in, int Lin, int stride, float* raw, float* stats, int last_only = 0,
                    const Gn* gn = nullptr) {
        if (err) return;
        auto it = a.conv.find(key);
        if (it == a.conv.end()) { err = "actor_tc_run: unknown conv"; return; }
        const ActorTcConv& cv = it->second;
        if (stride != cv.stride) { err = "actor_tc_run: conv stride differs from the packed one"; return; }
        const int Lout = (Lin - 1) / stride + 1;
        if (Lout % cv.fold || (last_only && cv.fold != 1)) { err = "actor_tc_run: fold does not divide the output length"; return; }
        const int Lf = Lout / cv.fold;                   // GEMM rows per actor
        int r_in = Lf >= 48 ? 16 : Lf / 3;               // 48->16, 24->8, 12->4, 6->2 : always 3 inner tiles
        if (gn) { r_in = 8; while (r_in < Lf) r_in <<= 1; }   // whole actors per tile: 6->8, 12->16, 24->32, 48->64 (3/4 of the rows used)
        const int r_out = 128 / r_in;
        const int C = cv.Cin_pad, N = cv.fold * cv.Cout;
        // k=3: window starts at padded row t*stride (original t*stride-1); k=1: padded row t*stride+1
        const int64_t base_off = (cv.ksize == 1) ? C : 0;
        const int64_t row_step = (int64_t)cv.fold * stride * C;
        alignas(64) unsigned char mh[128], ml[128];
        if ((err = tcg_encode_a(mh, in.hi + base_off, cv.Kpad, Lf, A, row_step, (int64_t)(Lin + 2) * C, r_in, r_out))) return;
        if ((err = tcg_encode_a(ml, in.lo + base_off, cv.Kpad, Lf, A, row_step, (int64_t)(Lin + 2) * C, r_in, r_out))) return;
        TcGemm g;
        g.amap_hi = mh; g.amap_lo = ml; g.wmap = cv.wmap; g.split = 1; g.k_blocks = cv.Kpad / 64;
        g.r_in = r_in; g.r_out = r_out; g.L_inner = Lf; g.n_outer = A;
        g.N = N; g.n_tile = N; g.C = raw; g.ldc = N; g.c_last_only = last_only; g.stats = stats; g.err = a.d_err;
        if (gn) {
            g.C = nullptr; g.stats = nullptr;
            g.gn_gamma = V(gn->key + ".weight"); g.gn_beta = V(gn->key + ".bias"); g.gn_C = cv.Cout;
            g.gn_inv_n = 1.f / (float)(Lout * cv.Cout); g.relu = gn->relu;
            g.gn_out_hi = gn->out.hi; g.gn_out_lo = gn->out.lo; g.gn_ld_group = (int64_t)(Lout + 2) * cv.Cout;
            if (gn->res) { g.gn_res_hi = gn->res->hi; g.gn_res_lo = gn->res->lo; }
            if (gn->res_raw) {
                g.gn_res_raw = gn->res_raw; g.gn_res_stats = gn->res_stats;
                g.gn_res_gamma = V(gn->res_key + ".weight"); g.gn_res_beta = V(gn->res_key + ".bias");
            }
            g.gn_up_prev = gn->up_prev; g.gn_out_f32 = gn->out_f32;
            if (err) return;
        }
        err = tcg_launch(g, sm_count, st);
    };
    auto apply = [&](const float* raw, const float* stats, const std::string& gk, int L, int C, int relu, HL out_hl, float* out_f32,
                     const float* res_raw = nullptr, const float* res_stats = nullptr, const std::string& rk = "",
                     const HL* res_hl = nullptr, const float* up_prev = nullptr, int last_only = 0) {
        if (err) return;
        TcApply q;
        q.up_prev = up_prev; q.last_only = last_only;
        q.raw = raw; q.stats = stats; q.gamma = V(gk + ".weight"); q.beta = V(gk + ".bias");
        if (res_raw) { q.res_raw = res_raw; q.res_stats = res_stats; q.res_gamma = V(rk + ".weight"); q.res_beta = V(rk + ".bias"); }
        if (res_hl) { q.res_hi = res_hl->hi; q.res_lo = res_hl->lo; }
        q.out_hi = out_hl.hi; q.out_lo = out_hl.lo; q.out_f32 = out_f32;
        q.A = A; q.L = L; q.C = C; q.relu = relu;
        if (!err) tcg_gn_apply(q, st);
    };
    const HL none{nullptr, nullptr};

    for (int i = 0; i < b.n_slack; ++i) cudaMemsetAsync(b.slack[i], 0, 1024 * sizeof(__half), st);
    tcg_actor_prep(actors, b.x0.hi, b.x0.lo, A, st);
    const int Cg[4] = {32, 64, 128, 256};
    HL cur = b.x0;
    int L = 48;
    for (int g = 0; g < 4 && !err; ++g) {
        char p[64];
        snprintf(p, sizeof p, "actor_net.groups.%d.", g);
        const std::string P(p);
        const int stride = g == 0 ? 1 : 2, Lo = (L - 1) / stride + 1, C = Cg[g];
        // block 0 (with conv shortcut): bn1 in conv1's epilogue; bn2 joins two GroupNorms, so it stays a separate pass
        if (a.gn_fused) {
            const Gn g1{P + "0.bn1", 1, b.t0, nullptr};
            conv(P + "0.conv1.weight", cur, L, stride, nullptr, nullptr, 0, &g1);
        } else {
            conv(P + "0.conv1.weight", cur, L, stride, b.raw1, b.st1);
            apply(b.raw1, b.st1, P + "0.bn1", Lo, C, 1, b.t0, nullptr);
        }
        conv(P + "0.downsample.0.weight", cur, L, stride, b.raw3, b.st3);
        if (a.gn_fused) {      // bn2 + GroupNorm of the down-sampled shortcut (its raw output and statistics) + ReLU in conv2's epilogue
            Gn g2{P + "0.bn2", 1, b.t1, nullptr};
            g2.res_raw = b.raw3; g2.res_stats = b.st3; g2.res_key = P + "0.downsample.1";
            conv(P + "0.conv2.weight", b.t0, Lo, 1, nullptr, nullptr, 0, &g2);
        } else {
            conv(P + "0.conv2.weight", b.t0, Lo, 1, b.raw2, b.st2);
            apply(b.raw2, b.st2, P + "0.bn2", Lo, C, 1, b.t1, nullptr, b.raw3, b.st3, P + "0.downsample.1");
        }
        // block 1 (identity shortcut): both GroupNorms in the epilogues
        if (a.gn_fused) {
            const Gn g1{P + "1.bn1", 1, b.t0, nullptr}, g2{P + "1.bn2", 1, b.o[g], &b.t1};
            conv(P + "1.conv1.weight", b.t1, Lo, 1, nullptr, nullptr, 0, &g1);
            conv(P + "1.conv2.weight", b.t0, Lo, 1, nullptr, nullptr, 0, &g2);
        } else {
            conv(P + "1.conv1.weight", b.t1, Lo, 1, b.raw1, b.st1);
            apply(b.raw1, b.st1, P + "1.bn1", Lo, C, 1, b.t0, nullptr);
            conv(P + "1.conv2.weight", b.t0, Lo, 1, b.raw2, b.st2);
            apply(b.raw2, b.st2, P + "1.bn2", Lo, C, 1, b.o[g], nullptr, nullptr, nullptr, "", &b.t1);
        }
        cur = b.o[g];
        L = Lo;
    }
    // FPN top-down
    const int Ls[4] = {48, 24, 12, 6};
    float* pyr = b.pyr0; float* nxt = b.pyr1;
    if (a.gn_fused) {           // lateral GroupNorm and the top-down step (:57-58) in the lateral conv's epilogue
        Gn g3{"actor_net.lateral.3.norm", 0, none, nullptr};
        g3.out_f32 = b.pyr0;
        conv("actor_net.lateral.3.conv.weight", b.o[3], 6, 1, nullptr, nullptr, 0, &g3);
        for (int i = 2; i >= 0 && !err; --i) {
            char p[64];
            snprintf(p, sizeof p, "actor_net.lateral.%d", i);
            // the finest level is only read as the (hi, lo) input / shortcut of the output block
            Gn gl{std::string(p) + ".norm", 0, i == 0 ? b.p : none, nullptr};
            gl.up_prev = pyr; gl.out_f32 = i == 0 ? nullptr : nxt;
            conv(std::string(p) + ".conv.weight", b.o[i], Ls[i], 1, nullptr, nullptr, 0, &gl);
            float* t = pyr; pyr = nxt; nxt = t;
        }
    } else {
        conv("actor_net.lateral.3.conv.weight", b.o[3], 6, 1, b.raw1, b.st1);
        apply(b.raw1, b.st1, "actor_net.lateral.3.norm", 6, 128, 0, none, b.pyr0);
        for (int i = 2; i >= 0 && !err; --i) {      // lateral GroupNorm and the top-down step (:57-58) in one pass over the conv output
            char p[64];
            snprintf(p, sizeof p, "actor_net.lateral.%d", i);
            conv(std::string(p) + ".conv.weight", b.o[i], Ls[i], 1, b.raw1, b.st1);
            apply(b.raw1, b.st1, std::string(p) + ".norm", Ls[i], 128, 0, i == 0 ? b.p : none, i == 0 ? nullptr : nxt,
                  nullptr, nullptr, "", nullptr, pyr);
            float* t = pyr; pyr = nxt; nxt = t;
        }
    }
    // output Res1d(128,128), identity shortcut = the pyramid top.  Only its last time step leaves ActorNet (:60), but both
    // GroupNorms take their statistics over all 48: conv2 runs in full, stores row 47 only, and is normalised there
    if (a.gn_fused) {
        const Gn g1{"actor_net.output.bn1", 1, b.t0, nullptr};
        conv("actor_net.output.conv1.weight", b.p, 48, 1, nullptr, nullptr, 0, &g1);
    } else {
        conv("actor_net.output.conv1.weight", b.p, 48, 1, b.raw1, b.st1);
        apply(b.raw1, b.st1, "actor_net.output.bn1", 48, 128, 1, b.t0, nullptr);
    }
    conv("actor_net.output.conv2.weight", b.t0, 48, 1, b.raw2, b.st2, 1);
    apply(b.raw2, b.st2, "actor_net.output.bn2", 48, 128, 1, none, out, nullptr, nullptr, "", &b.p, nullptr, 1);
    return err;
}

}  // namespace mind
