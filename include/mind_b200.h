/*
 * mind_b200.h -- C ABI of libmind_b200.so, the B200-native (sm_100a) implementation of
 * MIND's scenario-prediction hot path.
 *
 * The reference (HKUST-Aerial-Robotics/MIND) is pure Python and has no FFI of its own; the
 * boundary it does have is the string-imported network class
 *     net_cfg["network"] = "module:Class"      planners/mind/configs/networks/net_cfg.py:10
 * resolved in MINDPlanner.init_network           planners/mind/planner.py:42-49
 * and the two calls the tree generator makes on it
 *     network.pre_process(data); network(data_in)  planners/mind/scenario_tree.py:69-71
 * The Python class mind_b200.predictor:ScenePredNetB200 mirrors that surface and calls the
 * entry points below through ctypes (binding shown in INTEGRATION.md).  Every entry point
 * cites the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only (no torch types).  Device pointers are owned by
 * the caller; the library owns its packed weights and small descriptor tables.  All work is
 * enqueued asynchronously on the caller's stream; no hidden synchronisation in mind_forward /
 * mind_tree_* .  Every function returns 0 on success, non-zero on error with a message
 * available from mind_last_error().  There is no CPU fallback: without a CUDA device
 * mind_create fails.
 */
#ifndef MIND_B200_H
#define MIND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MindCtx MindCtx;

/* Network hyper-parameters are those of planners/mind/configs/networks/net_cfg.py:4-25
 * (fixed: 14 actor channels x 48 steps, 10 nodes x 16 lane channels, d=128, 6 fusion layers,
 * 8 heads, 6 modes, 60 predicted steps, Bezier order 7). */
#define MIND_D 128
#define MIND_ACTOR_C 14
#define MIND_ACTOR_T 48
#define MIND_LANE_NODES 10
#define MIND_LANE_C 16
#define MIND_MODES 6
#define MIND_PRED 60
#define MIND_RPE_C 5

/* precision modes of the N^2 contractions in the rela-fusion layers */
#define MIND_PREC_FP32 0 /* exact fp32 SIMT path: validation comparator, bit-stable ordering */
#define MIND_PREC_F16TC 1 /* tcgen05 kind::f16 operands, fp32 accumulate, fp16 edge stream */

/* ---- life cycle:  ScenePredNet(cfg, device)            planners/mind/planner.py:45 ---- */
int mind_create(MindCtx** out, int device);
void mind_destroy(MindCtx* ctx);
const char* mind_last_error(void);
/* library / build info string (arch, git-free) */
const char* mind_build_info(void);

/* ---- weights:  net.load_state_dict(ckpt["state_dict"])  planners/mind/planner.py:46-47 ----
 * Called once per state_dict entry with the reference's own key (328 keys) and a HOST fp32
 * pointer; mind_finalize_weights uploads and repacks (split of proj_memory 384 -> 3x128,
 * fused q/S/T projection, fp16 copies for the tensor-core path, transposed conv filters).
 * Two extra keys carry the Bezier bases the reference builds in __init__
 * (planners/mind/networks/network.py:449-464): "__bezier_T" [60*8], "__bezier_Tp" [60*7]. */
int mind_set_weight(MindCtx* ctx, const char* key, const float* host, int64_t numel);
int mind_finalize_weights(MindCtx* ctx);

/* options: "precision" (MIND_PREC_*), "chunk_scenes" (exact path workspace bound), "profile" (0/1), "graph" (0/1: replay a
 * captured CUDA graph when batch shape + pointers repeat), "tc_min_tokens" (MIND_PREC_F16TC only, default 128: scenes with
 * fewer tokens than this run the exact tier -- every N^2 contraction as a 3-term fp16 hi/lo tcgen05 product, fp32 edge --
 * instead of the fp16-operand fused kernel, whose operand rounding is not averaged out over a small scene's few keys;
 * 0 = fused kernel for every scene).  Diagnostic switches that select the un-fused comparator of a stage (each parity-tested
 * against the default): "lane_unfused", "node_unfused", "actor_gn_unfused" (GroupNorm as a separate pass behind every ActorNet
 * conv GEMM), "decoder_simt" (decoder linears on the fp32 SIMT GEMM), "actor_simt" (one-CTA-per-actor fp32 ActorNet). */
int mind_set_option(MindCtx* ctx, const char* name, int64_t value);

/* ---- one batched forward:  network(data_in)   planners/mind/networks/network.py:582-595 ----
 * Ragged batch of n_scenes scenes.  actor_off / lane_off are HOST prefix arrays [n_scenes+1].
 * Either rpe (HOST array of n_scenes DEVICE pointers, each [5, M_b, M_b] fp32, the 'scene'
 * entry of data['RPE'], planners/mind/utils.py:193-212) or ctrs+vecs (device [sum M_b, 2]
 * anchors; the library then evaluates get_rpe itself) must be given; M_b = Na_b + Nl_b.  */
typedef struct {
    int32_t n_scenes;
    const int32_t* actor_off; /* host [n_scenes+1] */
    const int32_t* lane_off;  /* host [n_scenes+1] */
    const float* actors;      /* dev [sumNa, 14, 48]   ACTORS    */
    const float* lanes;       /* dev [sumNl, 10, 16]   LANES     */
    const float* const* rpe;  /* host [n_scenes] of dev ptrs, or NULL */
    const float* ctrs;        /* dev [sumM, 2] or NULL (actors of scene b first, then lanes) */
    const float* vecs;        /* dev [sumM, 2] or NULL */
    const float* tgt_nodes;   /* dev [n_scenes, 10, 16] TGT_NODES */
    const float* tgt_rpe;     /* dev [n_scenes, 20]     TGT_RPE   */
} MindBatch;

/* Output layouts are the reference's (network.py:545-554), scenes concatenated along actors:
 *   cls [n_scenes, 6]            softmax mode probabilities (res_cls[b] = cls[b:b+1])
 *   reg [sumNa, 6, 60, 5]        x, y (actor-local), exp(cov) x3   (res_reg[b] = rows of b)
 *   vel [sumNa, 6, 60, 2]        res_aux[b][0]
 *   cov_vel [sumNa, 6, 60, 3]    res_aux[b][1]   (may be NULL)
 *   param [sumNa, 6, 8, 5]       res_aux[b][2] permuted to actor-major (may be NULL)      */
typedef struct {
    float* cls;
    float* reg;
    float* vel;
    float* cov_vel;
    float* param;
} MindOutputs;

/* bytes of caller-provided device workspace mind_forward needs for this batch shape */
int64_t mind_workspace_bytes(MindCtx* ctx, int32_t n_scenes, int32_t sum_actors, int32_t sum_lanes,
                             int32_t max_tokens /* max_b (Na_b+Nl_b+1) */);
/* same for one concrete batch (only n_scenes and the two offset arrays are read).  In the tensor-core mode the
 * requirement depends on the per-scene token counts: scenes below option "tc_min_tokens" take the exact tier (3-term
 * tcgen05 products, fp32 edge) on their own pair grid; mind_workspace_bytes() assumes every scene has max_tokens
 * tokens, which is exact for uniform batches (tree levels, the benchmark batch). */
int64_t mind_workspace_bytes_batch(MindCtx* ctx, const MindBatch* batch);

int mind_forward(MindCtx* ctx, const MindBatch* batch, const MindOutputs* out, void* workspace,
                 int64_t workspace_bytes, void* cuda_stream);

/* ---- batched host->device staging:  gpu(data[...])   planners/mind/utils.py:9-20, network.py:597-606 ----
 * The reference moves every per-scene tensor with its own .cuda() call (3 per scene: ACTOR_IDCS, LANE_IDCS,
 * RPE['scene']); at 256 scenes that is ~770 framework calls per step and the host becomes the bottleneck.
 * mind_upload_packed enqueues n host->device copies (host_ptrs[i], bytes[i]) into ONE caller-owned device
 * buffer in a single library call; copy i lands at dev_dst + offsets_out[i] (offsets are 256-byte aligned,
 * assigned in order; offsets_out is a HOST array of n entries).  Pinned sources copy asynchronously on
 * cuda_stream, pageable ones as the CUDA runtime stages them.  mind_upload_packed_bytes returns the
 * capacity needed for a given size list. */
int64_t mind_upload_packed_bytes(const int64_t* bytes, int32_t n);
int mind_upload_packed(const void* const* host_ptrs, const int64_t* bytes, int32_t n, void* dev_dst,
                       int64_t dst_capacity, int64_t* offsets_out, void* cuda_stream);

/* debug taps used by the stage-level parity tests: copy an internal stage buffer of the LAST
 * forward (still in the workspace) to a device buffer.  names: "actor_feat" [sumNa,128],
 * "lane_feat" [sumNl+B,128] (tgt polylines last), "actors_fused" [sumNa,128],
 * "cls_tok" [B,128].  Returns the number of floats written or <0. */
int64_t mind_debug_tap(MindCtx* ctx, const char* name, float* dst, int64_t capacity, void* cuda_stream);

/* ---- AIME scenario-tree step (reference planners/mind/scenario_tree.py) ---------------------
 * One call per depth level, between two batched network calls.  All pointers are DEVICE pointers
 * unless noted; F = frontier scenes of the level (they share n_actor: one tree = one root scene).
 *
 * mind_tree_level  replaces prune_merge (:281-412) + get_branch_time (:592-611) for the whole level:
 *   in : network outputs cls [F,6], reg [F*Na,6,60,5], vel [F*Na,6,60,2]; per scene ORIG [F,2],
 *        ROT [F,4] (row-major 2x2), actor anchors ctrs/vecs [F,Na,2]; parent histories (global
 *        frame) hpos/hvel [F,Na,50,2], hang/hcov [F,Na,50]; parent probability pprob [F], cur_t [F];
 *        target lane polyline tlane [n_tlane,2] (may be NULL), thresholds.
 *   out: per (scene, rank k): order [F,6] (mode index of rank k, descending probability),
 *        child histories cpos/cvel [F,6,Na,100,2], cang/ccov [F,6,Na,100], global predicted
 *        positions gpos [F,6,Na,60,2], cprob [F,6] = p_mode * p_parent, keep [F,6] (survives
 *        pruning + greedy topology merge), tb [F,6] (branch time; = pred_len when none).        */
typedef struct {
    int32_t n_frontier, n_actor, obs_len, pred_len, ego_idx, n_tlane;
    float tar_dist_thres;
    const float *cls, *reg, *vel;
    const float *orig, *rot, *ctrs, *vecs;
    const float *hpos, *hang, *hvel, *hcov;
    const float* pprob;
    const int32_t* cur_t;
    const float* tlane;
    float *cpos, *cang, *cvel, *ccov, *gpos;
    int32_t* order;
    float* cprob;
    int32_t *keep, *tb;
} MindTreeLevel;
int mind_tree_level(const MindTreeLevel* a, void* cuda_stream);

/* mind_tree_update  replaces update_obser (:467-567), get_new_lane_graph (utils.py:171-177),
 * get_high_level_command (:613-652) and the actor half of collate_fn (utils.py:114-139) for the
 * n_new children that branch: src [n_new,2] = (child row f*6+k, duration end_t-cur_t).
 *   out: new observation windows npos/nvel [n_new,Na,50,2], nang/ncov [n_new,Na,50]; norig, nrot,
 *        nctrs, nvecs; network inputs actors [n_new*Na,14,48], geometry geom_c/geom_v
 *        [n_new,Na+Nl,2] (the network evaluates get_rpe from these), tgt_nodes [n_new,10,16],
 *        tgt_rpe [n_new,20], tgt_pts [n_new,11,2].                                              */
typedef struct {
    int32_t n_new, n_actor, n_lane, n_tlane;
    float tar_time_ahead;
    const int32_t* src;
    const float *cpos, *cang, *cvel, *ccov;
    const float* ttype;                     /* [Na,50,7] TRAJS_TYPE of the ROOT observation, per step: the reference carries
                                             * it unchanged into every child scene (scenario_tree.py:486,524), including the
                                             * all-zero rows of steps at which the actor was not observed */
    const float *lane_ctrs, *lane_vecs;     /* [Nl,2] anchors of the stored lane graph */
    const float *tlane, *tinfo;             /* [n_tlane,2], [n_tlane,12] */
    float *npos, *nang, *nvel, *ncov, *norig, *nrot, *nctrs, *nvecs;
    float *actors, *geom_c, *geom_v, *tgt_nodes, *tgt_rpe, *tgt_pts;
} MindTreeUpdate;
int mind_tree_update(const MindTreeUpdate* u, void* cuda_stream);
const char* mind_tree_last_error(void);

/* ---- cost fields of the trajectory-tree optimiser (the step right after the scenario tree) -------
 * mind_cost_fields  replaces gen_dist_field (planners/ilqr/utils.py:5-22) and the per-node field assembly of
 * TrajectoryTreeOptimizer.init_warm_start_cost_tree / init_cost_tree (planners/mind/trajectory_tree.py:20-56, :58-124).
 * Grid: gx columns x gy rows of cell centres, cell (r, c) at (xs[c], ys[r]); the caller forms xs / ys with the
 * reference's own linspace + offset (utils.py:7-14) so that the coordinates are the same doubles.
 *   quad [gy,gx]            = (min over polyline segments of the clamped point-segment distance)^2
 *   fields [n_nodes,gy,gx]  = coef_tgt[n] * quad
 *                             + w_exo * sum_{e>=1} g(radius[n,e] - |p - mean[n,e]|),  g(f) = f + exo_cost_offset if f > 0 else 0
 *                             + w_ego * max(|p - mean[n,0]| - radius[n,0], 0)
 * n_actor = 0 gives the warm-start fields (first term only).  The caller forms coef_tgt = w_tgt * prob and
 * radius = cov + offset exactly as the reference does (fp32 sums) and passes them as fp64; node order = creation
 * order of the trajectory tree.  All arrays are DEVICE pointers, fp64 (the reference computes these in numpy fp64).
 * Asynchronous on the stream; error text from mind_cost_fields_last_error(). */
typedef struct {
    int32_t gx, gy;
    const double *xs, *ys;       /* [gx], [gy] cell-centre coordinates */
    int32_t n_lane_pts;
    const double* lane;          /* [n_lane_pts,2] target lane */
    int32_t n_nodes, n_actor;
    const double* coef_tgt;      /* [n_nodes] */
    const double* mean;          /* [n_nodes,n_actor,2], actor 0 = ego */
    const double* radius;        /* [n_nodes,n_actor] */
    double w_ego, w_exo, exo_cost_offset;
    double* quad;                /* out [gy,gx] */
    double* fields;              /* out [n_nodes,gy,gx] */
} MindCostFields;
int mind_cost_fields(const MindCostFields* a, void* cuda_stream);
const char* mind_cost_fields_last_error(void);

/* ---- tree iLQR of the trajectory-tree optimiser (HOST code, no device needed) ------------------------
 * mind_ilqr_tree_solve  replaces iLQR.fit (planners/ilqr/solver.py:80-167 with its forward rollout, recursive backward
 * pass, regularisation schedule and backtracking line search) on a TreeCost (planners/ilqr/cost.py:326-446) whose nodes
 * carry [PotentialField, StatePotential, StateConstraint] + [ControlPotential] (planners/ilqr/potential.py), for the
 * 6-state kinematic bicycle model of planners/mind/trajectory_tree.py:153-177 (state x, y, v, heading, a, steer; controls
 * da, dsteer).  Nodes are numbered in creation order; parent[0] = -1 (child of the root state x0), parent[i] < i.
 * All pointers are HOST pointers, fp64, row-major.  fields are the per-node cost fields (mind_cost_fields output copied
 * to the host, or any other source) on the grid xs_grid [gx] x ys_grid [gy] with cell size res and offset field_offset.
 * Returns 0 and writes the optimised state / control paths; error text from mind_ilqr_last_error(). */
typedef struct {
    int32_t n_nodes;
    const int32_t* parent;       /* [n] */
    const double* x0;            /* [6] */
    double dt, wheelbase;
    int32_t gx, gy;
    double res;
    const double* field_offset;  /* [2] */
    const double *xs_grid, *ys_grid;
    const double* fields;        /* [n,gy,gx] */
    const double* w_state;       /* [n,6,6]  StatePotential weight */
    const double* des_state;     /* [n,6] */
    const double* w_con;         /* [n,6,6]  StateConstraint weight */
    const double *lower, *upper; /* [6] */
    const double* w_ctrl;        /* [n,2,2]  ControlPotential weight */
    int32_t max_iter;            /* <= 0: 100 (solver.py:80) */
    const double* us_init;       /* [n,2] */
    double *xs_out, *us_out;     /* [n,6], [n,2] */
    int32_t* iterations;         /* out, may be NULL */
    double* cost;                /* out: trajectory cost of the last forward rollout, may be NULL */
} MindIlqrTree;
int mind_ilqr_tree_solve(const MindIlqrTree* p);
const char* mind_ilqr_last_error(void);
/* diagnostic: PotentialField.get_potential / get_gradient / get_hessian (potential.py:71-104) of node `node` at (x, y);
 * only the grid members of p are read.  out6 = {value, d/dx, d/dy, d2/dx2, d2/dxdy, d2/dy2}. */
int mind_debug_field_eval(const MindIlqrTree* p, int32_t node, double x, double y, double* out6);

/* bring-up self test of the TMA + tcgen05 + TMEM plumbing: D[0:128*128] = A . W^T (fp16 operands,
 * fp32 accumulate), D[128*128: 2*128*128] = the A tile read back through the software swizzle,
 * D[2*128*128: 3*128*128] = the same product with the A operand staged in tensor memory.
 * All three are HOST buffers (A, W: 128*128 floats; D: 3*128*128 floats). */
int mind_tc_selftest(const float* A_host, const float* W_host, float* D_host);

/* Host-only diagnostic (no device needed): the static schedule the fused rela-fusion layer would use for scenes
 * of n_tokens[b] = actors + lanes + 1 tokens on sm_count CTAs.  work receives 8 int32 per entry
 * {b, j0, n, ch0, ch1, slot, mode, 0}: first `grid` range headers {first item, one past the last item}, then the work
 * items (mode 0: queries j0..j0+15, key chunks [ch0, ch1) of 8 keys; mode 1: query j0 only, chunks of 128 keys;
 * slot >= 0: one part of a key-split item).  info receives {n_entries, grid, n_merge_jobs, n_slots}.  Returns 0, or
 * -1 when capacity (in entries) is too small (info[0] then holds the required count). */
int mind_debug_fusion_schedule(const int32_t* n_tokens, int32_t n_scenes, int32_t sm_count, int32_t* work,
                               int32_t capacity, int32_t* info);

/* Host-only diagnostic (no device needed): the parameter tables the fp16 edge-init kernel is driven by, from the four
 * tensors of fusion_net.proj_rpe_scene (Linear(5,128) + LayerNorm(128) + ReLU, planners/mind/networks/network.py:282-286,
 * 326-330).  quad21 = {c0, L[5], Q upper triangle by rows with doubled off-diagonals}: LayerNorm variance of the
 * projection as a quadratic form of the 5 RPE values; tab896 = [32 lanes][2 channel pairs][7][2]: gamma * centred W (5),
 * gamma * centred b, beta of channels 4*lane + 2*pair + {0,1}.  With rstd = rsqrt(var + 1e-5) the layer's pre-ReLU output
 * is  y[c] = sum_k tab[c][k] * rstd * r[k] + tab[c][5] * rstd + tab[c][6]. */
int mind_debug_edge_init_pack(const float* W, const float* b, const float* gamma, const float* beta, float* tab896,
                              float* quad21);

/* Host-only diagnostic: the GEMM operand the ActorNet conv engine uses for a Conv1d weight [Cout][Cin][ks] (padding ks/2,
 * planners/mind/networks/layers.py:36-60) when `fold` consecutive output steps share one GEMM row: out receives
 * [fold*Cout][Kpad] fp32 (capacity in floats; Kpad is returned, or -1 if capacity is too small / arguments are bad).
 * GEMM row p, column u*Cout + o = sum over the row's window (padded input rows fold*p*stride ..., Cin_pad channels each) =
 * output channel o at step fold*p + u. */
int mind_debug_conv_fold_pack(const float* w, int32_t Cout, int32_t Cin, int32_t Cin_pad, int32_t ksize, int32_t stride,
                              int32_t fold, float* out, int64_t capacity);

/* cudaDeviceSynchronize + kernel-side protocol error flag (0 = clean) */
int mind_sync_check(MindCtx* ctx);

/* per-stage device timing: with option "profile"=1 every forward records CUDA events around its
 * stages on the caller's stream; this call synchronises and drains them as text lines
 * "tag total_ms count" (tags: actor_net lane_net tokens edge_init node_pre fusion_tc
 * fusion_tc_last node_post fusion_other decoder). */
int mind_profile_read(MindCtx* ctx, char* buf, int64_t capacity);

/* number of kernels the library launched since creation (bench.py's gpu_launches) */
int64_t mind_launch_count(MindCtx* ctx);

/* option "graph" = 1: a forward whose batch shape AND pointer set (inputs, outputs, workspace, stream) repeat is
 * captured into a CUDA graph on its second appearance and replayed afterwards (one graph launch instead of ~170
 * kernel launches; descriptor tables are private to the graph).  Meant for the scenario tree's small level batches
 * with persistent buffers.  mind_graph_replays counts the forwards served by a replay. */
int64_t mind_graph_replays(MindCtx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MIND_B200_H */
