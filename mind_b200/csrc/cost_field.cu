// Cost fields of the trajectory-tree optimiser (reference planners/mind/trajectory_tree.py:20-124,
// planners/ilqr/utils.py:5-22, common/geometry.py:27-30,70-78): the step right after the scenario tree.
//   k_lane_dist_sq   squared distance of every grid cell centre to the target-lane polyline (min over segments of the
//                    clamped point-segment distance), once per plan call
//   k_node_fields    one [gy, gx] field per trajectory-tree node:
//                    coef * d_lane^2 + w_exo * sum_exo g(r_e - |p - mu_e|) + w_ego * max(|p - mu_0| - r_0, 0),
//                    g(f) = max(f, 0) (+ cost offset where positive); exo actors accumulated in index order like the
//                    reference's loop.  fp64 throughout (the reference works in numpy fp64).
// Both kernels are streaming writes: 8 B per cell per node (HBM-bound, 0.5 MB per node at 256 x 256); the polyline and
// a node's actor table sit in shared memory.
#include "../../include/mind_b200.h"
#include "kernels.h"
#include <math.h>
#include <cstdio>

namespace mind {

constexpr int kCfThreads = 256;

__global__ void __launch_bounds__(kCfThreads) k_lane_dist_sq(MindCostFields a) {
    extern __shared__ double s_lane[];                       // [n_lane_pts][2]
    for (int i = threadIdx.x; i < a.n_lane_pts * 2; i += blockDim.x) s_lane[i] = a.lane[i];
    __syncthreads();
    const int64_t cells = (int64_t)a.gx * a.gy;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < cells; p += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(p / a.gx), c = (int)(p % a.gx);
        const double x = a.xs[c], y = a.ys[r];               // cell centres exactly as the caller's linspace formed them
        double best = INFINITY;
        for (int j = 0; j + 1 < a.n_lane_pts; ++j) {
            const double ax = s_lane[2 * j], ay = s_lane[2 * j + 1];
            const double lx = s_lane[2 * j + 2] - ax, ly = s_lane[2 * j + 3] - ay;
            double t = ((x - ax) * lx + (y - ay) * ly) / (lx * lx + ly * ly);
            t = fmin(fmax(t, 0.0), 1.0);
            const double dx = x - (ax + t * lx), dy = y - (ay + t * ly);
            best = fmin(best, sqrt(dx * dx + dy * dy));
        }
        a.quad[p] = best * best;
    }
}

__global__ void __launch_bounds__(kCfThreads) k_node_fields(MindCostFields a) {
    extern __shared__ double s_act[];                        // [n_actor][3]: mu_x, mu_y, radius
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < a.n_actor; i += blockDim.x) {
        s_act[3 * i] = a.mean[((int64_t)n * a.n_actor + i) * 2];
        s_act[3 * i + 1] = a.mean[((int64_t)n * a.n_actor + i) * 2 + 1];
        s_act[3 * i + 2] = a.radius[(int64_t)n * a.n_actor + i];
    }
    __syncthreads();
    const int64_t cells = (int64_t)a.gx * a.gy;
    const double coef = a.coef_tgt[n];
    double* out = a.fields + (int64_t)n * cells;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < cells; p += (int64_t)gridDim.x * blockDim.x) {
        double v = coef * a.quad[p];
        if (a.n_actor > 0) {
            const int r = (int)(p / a.gx), c = (int)(p % a.gx);
            const double x = a.xs[c], y = a.ys[r];
            double acc = 0.0;
            for (int e = 1; e < a.n_actor; ++e) {
                const double dx = x - s_act[3 * e], dy = y - s_act[3 * e + 1];
                const double f = s_act[3 * e + 2] - sqrt(dx * dx + dy * dy);
                if (f > 0.0) acc += f + a.exo_cost_offset;
            }
            const double dx = x - s_act[0], dy = y - s_act[1];
            const double ego = fmax(sqrt(dx * dx + dy * dy) - s_act[2], 0.0);
            v = v + a.w_exo * acc + a.w_ego * ego;
        }
        out[p] = v;
    }
}

}  // namespace mind

static thread_local char g_cferr[256] = "";
extern "C" const char* mind_cost_fields_last_error(void) { return g_cferr; }

extern "C" int mind_cost_fields(const MindCostFields* a, void* cuda_stream) {
    using namespace mind;
    if (!a || a->gx <= 0 || a->gy <= 0 || !a->xs || !a->ys || a->n_lane_pts < 2 || a->n_nodes < 0 || a->n_actor < 0 || !a->lane || !a->quad ||
        (a->n_nodes > 0 && (!a->fields || !a->coef_tgt)) || (a->n_nodes > 0 && a->n_actor > 0 && (!a->mean || !a->radius))) {
        snprintf(g_cferr, sizeof g_cferr, "mind_cost_fields: bad argument");
        return 1;
    }
    const size_t lane_smem = sizeof(double) * 2 * (size_t)a->n_lane_pts, act_smem = sizeof(double) * 3 * (size_t)a->n_actor;
    if (lane_smem > 200 * 1024 || act_smem > 48 * 1024) {
        snprintf(g_cferr, sizeof g_cferr, "mind_cost_fields: polyline (%d points) or actor table (%d) too large", a->n_lane_pts, a->n_actor);
        return 1;
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int64_t cells = (int64_t)a->gx * a->gy;
    const int blocks = (int)((cells + kCfThreads - 1) / kCfThreads);
    if (lane_smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_lane_dist_sq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lane_smem) != cudaSuccess) {
        snprintf(g_cferr, sizeof g_cferr, "mind_cost_fields: cudaFuncSetAttribute failed");
        return 1;
    }
    k_lane_dist_sq<<<blocks, kCfThreads, lane_smem, st>>>(*a);
    ++g_launches;
    if (a->n_nodes > 0) {
        // a node's cells over at most 148 x 4 CTAs per node row; rows = nodes
        const int bx = blocks < 592 ? blocks : 592;
        k_node_fields<<<dim3((unsigned)bx, (unsigned)a->n_nodes), kCfThreads, act_smem, st>>>(*a);
        ++g_launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_cferr, sizeof g_cferr, "mind_cost_fields: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
