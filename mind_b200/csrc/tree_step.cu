// AIME scenario-tree step kernels (reference planners/mind/scenario_tree.py): everything the tree
// does between two batched network calls, flattened over the whole frontier of one depth level.
//   k_tree_expand  (:314-348)  mode sort + actor-local -> scene -> global transform of the K=6
//                              predictions of every frontier scene, covariance accumulation,
//                              history concatenation  [F,6,Na,100,.]
//   k_tree_select  (:369-410, :592-611)  probability / target-lane pruning, topology signature,
//                              greedy merge, branch-time scan
//   k_tree_update  (:467-567, :613-652, utils.py:171-212)  slide the observation window to the
//                              branch time, re-centre on the ego, per-actor re-normalisation,
//                              actor features, lane anchors, high-level command -> next level's
//                              network inputs (RPE is evaluated inside the network from anchors)
// All frontier scenes of one tree share Na (they descend from one root scene).
#include "../../include/mind_b200.h"
#include "kernels.h"
#include <math.h>
#include <algorithm>
#include <cstdio>

namespace mind {

#define PI_F 3.14159265358979323846f

__device__ __forceinline__ float wrap_angle(float a) { return atan2f(sinf(a), cosf(a)); }

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tree_expand(MindTreeLevel a) {
    // block = (f, rank k, actor i); threads over the 110 concatenated steps
    const int blk = blockIdx.x;
    const int i = blk % a.n_actor;
    const int k = (blk / a.n_actor) % 6;
    const int f = blk / (a.n_actor * 6);
    __shared__ int s_mode;
    if (threadIdx.x == 0) {
        // descending argsort of the 6 mode probabilities: mode whose rank is k
        const float* c = a.cls + f * 6;
        int mode = 0;
        for (int m = 0; m < 6; ++m) {
            int rank = 0;
            for (int m2 = 0; m2 < 6; ++m2) rank += (c[m2] > c[m] || (c[m2] == c[m] && m2 < m)) ? 1 : 0;
            if (rank == k) mode = m;
        }
        s_mode = mode;
        if (i == 0) a.order[f * 6 + k] = mode;
    }
    __syncthreads();
    const int m = s_mode;
    const int64_t fa = (int64_t)f * a.n_actor + i;                  // row in parent arrays / net outputs
    const int64_t ck = ((int64_t)(f * 6 + k) * a.n_actor + i);      // row in child arrays
    const float vx = a.vecs[fa * 2], vy = a.vecs[fa * 2 + 1];
    const float th = atan2f(vy, vx);
    const float c = cosf(th), s = sinf(th);
    const float R00 = a.rot[f * 4 + 0], R01 = a.rot[f * 4 + 1], R10 = a.rot[f * 4 + 2], R11 = a.rot[f * 4 + 3];
    const float thg = atan2f(R10, R00);
    const float ox = a.orig[f * 2], oy = a.orig[f * 2 + 1];
    const float cx = a.ctrs[fa * 2], cy = a.ctrs[fa * 2 + 1];
    const float cov_last = a.hcov[fa * 50 + 49];
    for (int t = threadIdx.x; t < 110; t += blockDim.x) {
        if (t < 50) {
            a.cpos[(ck * 100 + t) * 2] = a.hpos[(fa * 50 + t) * 2];
            a.cpos[(ck * 100 + t) * 2 + 1] = a.hpos[(fa * 50 + t) * 2 + 1];
            a.cvel[(ck * 100 + t) * 2] = a.hvel[(fa * 50 + t) * 2];
            a.cvel[(ck * 100 + t) * 2 + 1] = a.hvel[(fa * 50 + t) * 2 + 1];
            a.cang[ck * 100 + t] = a.hang[fa * 50 + t];
            a.ccov[ck * 100 + t] = a.hcov[fa * 50 + t];
        } else {
            const int p = t - 50;
            const float* r = a.reg + ((fa * 6 + m) * 60 + p) * 5;
            const float* v = a.vel + ((fa * 6 + m) * 60 + p) * 2;
            // actor-local -> scene: p . R_a^T + ctr ;  scene -> global: p . ROT^T + ORIG  (:333-339)
            const float sx = r[0] * c - r[1] * s + cx, sy = r[0] * s + r[1] * c + cy;
            const float gx = sx * R00 + sy * R01 + ox, gy = sx * R10 + sy * R11 + oy;
            const float svx = v[0] * c - v[1] * s, svy = v[0] * s + v[1] * c;
            const float gvx = svx * R00 + svy * R01, gvy = svx * R10 + svy * R11;
            a.gpos[(ck * 60 + p) * 2] = gx;
            a.gpos[(ck * 60 + p) * 2 + 1] = gy;
            if (t < 100) {
                a.cpos[(ck * 100 + t) * 2] = gx; a.cpos[(ck * 100 + t) * 2 + 1] = gy;
                a.cvel[(ck * 100 + t) * 2] = gvx; a.cvel[(ck * 100 + t) * 2 + 1] = gvy;
                a.cang[ck * 100 + t] = atan2f(v[1], v[0]) + th + thg;                 // (:311,341)
                a.ccov[ck * 100 + t] = fmaxf(r[2], r[3]) + cov_last;                  // (:325,343)
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tree_select(MindTreeLevel a) {
    const int f = blockIdx.x;
    extern __shared__ float sh[];
    float* topo = sh;                                   // [6][n_actor-1]
    __shared__ float s_red[128];
    __shared__ int s_alive[6], s_keep[6];
    __shared__ float s_prob[6];
    const int ne = a.n_actor - 1;
    // topology signature: cumulative wrapped angle change of (exo - ego) over the 60 predicted steps (:382-392)
    for (int idx = threadIdx.x; idx < 6 * ne; idx += blockDim.x) {
        const int k = idx / ne, e = idx % ne + 1;
        const int64_t base = (int64_t)(f * 6 + k) * a.n_actor;
        const float* pe = a.gpos + (base + e) * 120;
        const float* p0 = a.gpos + (base + 0) * 120;
        float prev = 0.f, sum = 0.f;
        for (int t = 0; t < 60; ++t) {
            float dx = pe[t * 2] - p0[t * 2], dy = pe[t * 2 + 1] - p0[t * 2 + 1];
            const float n = sqrtf(dx * dx + dy * dy);
            dx /= n; dy /= n;
            const float ang = atan2f(dy, dx);
            if (t > 0) sum += wrap_angle(ang - prev);
            prev = ang;
        }
        topo[k * ne + (e - 1)] = sum;
    }
    // pruning: probability (:369) and distance of the ego end point to the target lane (:373-379)
    for (int k = 0; k < 6; ++k) {
        const int m = a.order[f * 6 + k];
        const float prob = a.cls[f * 6 + m] * a.pprob[f];
        int alive = !(prob < 0.001f);
        if (a.tlane && a.ego_idx >= 0) {
            const int64_t ce = ((int64_t)(f * 6 + k) * a.n_actor + a.ego_idx) * 100 + 99;
            const float px = a.cpos[ce * 2], py = a.cpos[ce * 2 + 1];
            float best = INFINITY;
            for (int sgm = threadIdx.x; sgm < a.n_tlane - 1; sgm += blockDim.x) {
                const float x1 = a.tlane[sgm * 2], y1 = a.tlane[sgm * 2 + 1];
                const float sx = a.tlane[sgm * 2 + 2] - x1, sy = a.tlane[sgm * 2 + 3] - y1;
                float tt = ((px - x1) * sx + (py - y1) * sy) / (sx * sx + sy * sy);
                tt = fminf(fmaxf(tt, 0.f), 1.f);
                const float qx = x1 + tt * sx - px, qy = y1 + tt * sy - py;
                best = fminf(best, sqrtf(qx * qx + qy * qy));
            }
            s_red[threadIdx.x] = best;
            __syncthreads();
            for (int o = 64; o > 0; o >>= 1) {
                if (threadIdx.x < o) s_red[threadIdx.x] = fminf(s_red[threadIdx.x], s_red[threadIdx.x + o]);
                __syncthreads();
            }
            if (s_red[0] - a.ccov[ce] > a.tar_dist_thres) alive = 0;
            __syncthreads();
        }
        if (threadIdx.x == 0) { s_alive[k] = alive; s_prob[k] = prob; }
    }
    __syncthreads();
    // greedy merge in rank order (:396-410): a candidate is kept iff it differs from every kept one
    if (threadIdx.x == 0) {
        for (int k = 0; k < 6; ++k) {
            int keep = s_alive[k];
            for (int s = 0; s < k && keep; ++s) {
                if (!s_keep[s]) continue;
                int differs = 0;
                for (int e = 0; e < ne; ++e)
                    if (fabsf(wrap_angle(topo[s * ne + e] - topo[k * ne + e])) - PI_F / 6.f > 0.f) { differs = 1; break; }
                if (!differs) keep = 0;
            }
            s_keep[k] = keep;
        }
    }
    __syncthreads();
    // branch time scan (:592-611) for the kept children
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        const int cur_t = a.cur_t[f], end_t = a.pred_len;
        int tb = end_t;
        if (s_keep[k]) {
            const int cmp_t = a.obs_len + cur_t + (cur_t == 0 ? 1 : 0);
            for (int t = cur_t + 1; t < end_t && tb == end_t; ++t) {
                if (t & 1) continue;
                for (int i = 0; i < a.n_actor; ++i) {
                    const float* cv = a.ccov + ((int64_t)(f * 6 + k) * a.n_actor + i) * 100;
                    if (cv[a.obs_len + t] / cv[cmp_t] > 9.f) { tb = t; break; }
                }
            }
        }
        a.keep[f * 6 + k] = s_keep[k];
        a.cprob[f * 6 + k] = s_prob[k];
        a.tb[f * 6 + k] = tb;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tree_update(MindTreeUpdate u) {
    const int g = blockIdx.x;
    const int src = u.src[g * 2], dur = u.src[g * 2 + 1];          // child row (f*6+k), end_t - cur_t
    const int Na = u.n_actor;
    __shared__ float s_f[8];                                       // orig(2) c s theta | ego speed
    const int64_t cb = (int64_t)src * Na;
    // 1. window [dur, dur+50) of the child history becomes the new observation (:471-482)
    for (int idx = threadIdx.x; idx < Na * 50; idx += blockDim.x) {
        const int i = idx / 50, t = idx % 50;
        const int64_t s = (cb + i) * 100 + dur + t, d = ((int64_t)g * Na + i) * 50 + t;
        u.npos[d * 2] = u.cpos[s * 2]; u.npos[d * 2 + 1] = u.cpos[s * 2 + 1];
        u.nvel[d * 2] = u.cvel[s * 2]; u.nvel[d * 2 + 1] = u.cvel[s * 2 + 1];
        u.nang[d] = u.cang[s];
        u.ncov[d] = u.ccov[s];
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                        // ego pose at the last observed step (:496)
        const int64_t e = ((int64_t)g * Na + 0) * 50 + 49;
        const float th = u.nang[e];
        s_f[0] = u.npos[e * 2]; s_f[1] = u.npos[e * 2 + 1]; s_f[2] = cosf(th); s_f[3] = sinf(th); s_f[4] = th;
        u.norig[g * 2] = s_f[0]; u.norig[g * 2 + 1] = s_f[1];
        u.nrot[g * 4 + 0] = s_f[2]; u.nrot[g * 4 + 1] = -s_f[3]; u.nrot[g * 4 + 2] = s_f[3]; u.nrot[g * 4 + 3] = s_f[2];
    }
    __syncthreads();
    const float ox = s_f[0], oy = s_f[1], C = s_f[2], S = s_f[3], TH = s_f[4];
    const int M = Na + u.n_lane;
    // 2. per actor: scene frame (p - orig).ROT, then its own frame at step 49; features [14,48] (:498-529, utils.py:114-132)
    for (int i = threadIdx.x; i < Na; i += blockDim.x) {
        const int64_t r = ((int64_t)g * Na + i) * 50;
        const float lx = u.npos[(r + 49) * 2] - ox, ly = u.npos[(r + 49) * 2 + 1] - oy;
        const float cxs = lx * C + ly * S, cys = -lx * S + ly * C;         // actor anchor in the scene frame
        const float tha = u.nang[r + 49] - TH;
        const float ca = cosf(tha), sa = sinf(tha);
        u.nctrs[((int64_t)g * Na + i) * 2] = cxs; u.nctrs[((int64_t)g * Na + i) * 2 + 1] = cys;
        u.nvecs[((int64_t)g * Na + i) * 2] = ca; u.nvecs[((int64_t)g * Na + i) * 2 + 1] = sa;
        u.geom_c[((int64_t)g * M + i) * 2] = cxs; u.geom_c[((int64_t)g * M + i) * 2 + 1] = cys;
        u.geom_v[((int64_t)g * M + i) * 2] = ca; u.geom_v[((int64_t)g * M + i) * 2 + 1] = sa;
        float* feat = u.actors + ((int64_t)g * Na + i) * 14 * 48;
        float ppx = 0.f, ppy = 0.f;
        for (int t = 0; t < 50; ++t) {
            const float gx = u.npos[(r + t) * 2] - ox, gy = u.npos[(r + t) * 2 + 1] - oy;
            const float sx = gx * C + gy * S - cxs, sy = -gx * S + gy * C - cys;
            const float px = sx * ca + sy * sa, py = -sx * sa + sy * ca;       // (p_scene - o_i) . rot_i
            const float gvx = u.nvel[(r + t) * 2], gvy = u.nvel[(r + t) * 2 + 1];
            const float svx = gvx * C + gvy * S, svy = -gvx * S + gvy * C;
            const float vx = svx * ca + svy * sa, vy = -svx * sa + svy * ca;
            const float an = u.nang[r + t] - TH - tha;
            if (i == 0 && t == 49) s_f[5] = sqrtf(vx * vx + vy * vy);
            if (t >= 2) {
                const int tt = t - 2;
                feat[0 * 48 + tt] = px - ppx; feat[1 * 48 + tt] = py - ppy;     // displacement (first step would be 0)
                feat[2 * 48 + tt] = cosf(an); feat[3 * 48 + tt] = sinf(an);
                feat[4 * 48 + tt] = vx; feat[5 * 48 + tt] = vy;
                for (int c7 = 0; c7 < 7; ++c7) feat[(6 + c7) * 48 + tt] = u.ttype[((int64_t)i * 50 + t) * 7 + c7];   // root's per-step rows (:486,524)
                feat[13 * 48 + tt] = 1.f;
            }
            ppx = px; ppy = py;
        }
    }
    // 3. lane anchors (utils.py:171-177; the reference transforms the ROOT-frame anchors with the new global pose)
    for (int l = threadIdx.x; l < u.n_lane; l += blockDim.x) {
        const float lx = u.lane_ctrs[l * 2] - ox, ly = u.lane_ctrs[l * 2 + 1] - oy;
        const float vx = u.lane_vecs[l * 2], vy = u.lane_vecs[l * 2 + 1];
        u.geom_c[((int64_t)g * M + Na + l) * 2] = lx * C + ly * S; u.geom_c[((int64_t)g * M + Na + l) * 2 + 1] = -lx * S + ly * C;
        u.geom_v[((int64_t)g * M + Na + l) * 2] = vx * C + vy * S; u.geom_v[((int64_t)g * M + Na + l) * 2 + 1] = -vx * S + vy * C;
    }
    __syncthreads();
    // 4. high-level command (:613-652) + target RPE (utils.py:193-212 on the 2-point set {anchor, ego})
    if (threadIdx.x == 0) {
        const int n = u.n_tlane;
        int closest = 0; float best = INFINITY;
        for (int p = 0; p < n; ++p) {
            const float dx = u.tlane[p * 2] - ox, dy = u.tlane[p * 2 + 1] - oy;
            const float d = sqrtf(dx * dx + dy * dy);
            if (d < best) { best = d; closest = p; }
        }
        float travel = fmaxf(s_f[5], 0.5f) * u.tar_time_ahead;
        int idx = closest;
        while (idx < n - 1 && travel > 0.f) {
            ++idx;
            const float dx = u.tlane[idx * 2] - u.tlane[idx * 2 - 2], dy = u.tlane[idx * 2 + 1] - u.tlane[idx * 2 - 1];
            travel -= sqrtf(dx * dx + dy * dy);
        }
        if (idx == n - 1) --idx;
        idx = max(5, min(idx, n - 6));
        float lx[11], ly[11], mx = 0.f, my = 0.f;
        for (int p = 0; p < 11; ++p) {
            const float gx = u.tlane[(idx - 5 + p) * 2], gy = u.tlane[(idx - 5 + p) * 2 + 1];
            u.tgt_pts[(g * 11 + p) * 2] = gx; u.tgt_pts[(g * 11 + p) * 2 + 1] = gy;
            lx[p] = (gx - ox) * C + (gy - oy) * S; ly[p] = -(gx - ox) * S + (gy - oy) * C;
            mx += lx[p]; my += ly[p];
        }
        mx /= 11.f; my /= 11.f;
        float ax = lx[10] - lx[0], ay = ly[10] - ly[0];
        const float an = sqrtf(ax * ax + ay * ay);
        ax /= an; ay /= an;
        for (int p = 0; p < 11; ++p) {                         // to the instance frame: (p - anchor) . [[ax,-ay],[ay,ax]]
            const float qx = lx[p] - mx, qy = ly[p] - my;
            lx[p] = qx * ax + qy * ay; ly[p] = -qx * ay + qy * ax;
        }
        float* tn = u.tgt_nodes + (int64_t)g * 160;
        for (int p = 0; p < 10; ++p) {
            tn[p * 16 + 0] = (lx[p] + lx[p + 1]) / 2.f; tn[p * 16 + 1] = (ly[p] + ly[p + 1]) / 2.f;
            tn[p * 16 + 2] = lx[p + 1] - lx[p]; tn[p * 16 + 3] = ly[p + 1] - ly[p];
            for (int c = 0; c < 12; ++c) tn[p * 16 + 4 + c] = u.tinfo[(idx - 5 + p + 1) * 12 + c];
        }
        // TGT_RPE [5][2][2]: points {anchor, ego anchor}; entry [a][b] relates vec[b] to vec[a] and ctr[b]-ctr[a]
        const float pc[2][2] = {{mx, my}, {u.nctrs[(int64_t)g * Na * 2], u.nctrs[(int64_t)g * Na * 2 + 1]}};
        const float pv[2][2] = {{ax, ay}, {u.nvecs[(int64_t)g * Na * 2], u.nvecs[(int64_t)g * Na * 2 + 1]}};
        float* tr = u.tgt_rpe + (int64_t)g * 20;
        for (int aa = 0; aa < 2; ++aa)
            for (int bb = 0; bb < 2; ++bb) {
                const float dx = pc[bb][0] - pc[aa][0], dy = pc[bb][1] - pc[aa][1];
                const float dist = sqrtf(dx * dx + dy * dy);
                const float nb = sqrtf(pv[bb][0] * pv[bb][0] + pv[bb][1] * pv[bb][1]);
                const float na = sqrtf(pv[aa][0] * pv[aa][0] + pv[aa][1] * pv[aa][1]);
                const float d1 = nb * na + 1e-10f, d2 = nb * dist + 1e-10f;
                tr[0 * 4 + aa * 2 + bb] = (pv[bb][0] * pv[aa][0] + pv[bb][1] * pv[aa][1]) / d1;
                tr[1 * 4 + aa * 2 + bb] = (pv[bb][0] * pv[aa][1] - pv[bb][1] * pv[aa][0]) / d1;
                tr[2 * 4 + aa * 2 + bb] = (pv[bb][0] * dx + pv[bb][1] * dy) / d2;
                tr[3 * 4 + aa * 2 + bb] = (pv[bb][0] * dy - pv[bb][1] * dx) / d2;
                tr[4 * 4 + aa * 2 + bb] = dist * 2.f / 100.f;
            }
    }
}

}  // namespace mind

static thread_local char g_terr[256] = "";
extern "C" const char* mind_tree_last_error(void) { return g_terr; }

extern "C" int mind_tree_level(const MindTreeLevel* a, void* cuda_stream) {
    using namespace mind;
    if (!a || a->n_frontier <= 0 || a->n_actor <= 0 || a->n_actor > 256) { snprintf(g_terr, sizeof g_terr, "mind_tree_level: bad argument"); return 1; }
    // child histories are 100-step rows (50 observed + 50 predicted, planners/mind/planner.py:20-21,53): the kernels index them so
    if (a->obs_len != 50 || a->pred_len != 50) {
        snprintf(g_terr, sizeof g_terr, "mind_tree_level: obs_len / pred_len must be 50 / 50 (got %d / %d)", a->obs_len, a->pred_len);
        return 1;
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    k_tree_expand<<<a->n_frontier * 6 * a->n_actor, 128, 0, st>>>(*a);
    k_tree_select<<<a->n_frontier, 128, sizeof(float) * 6 * (size_t)std::max(a->n_actor - 1, 1), st>>>(*a);
    g_launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_terr, sizeof g_terr, "mind_tree_level: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

extern "C" int mind_tree_update(const MindTreeUpdate* u, void* cuda_stream) {
    using namespace mind;
    if (!u || u->n_new <= 0) { snprintf(g_terr, sizeof g_terr, "mind_tree_update: bad argument"); return 1; }
    k_tree_update<<<u->n_new, 128, 0, (cudaStream_t)cuda_stream>>>(*u);
    g_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_terr, sizeof g_terr, "mind_tree_update: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
