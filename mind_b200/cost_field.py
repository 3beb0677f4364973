"""Cost fields of MIND's trajectory-tree optimiser on the GPU (the step right after the scenario tree).

`cost_fields(scen_tree, x0, target_lane, cfg, device, warm)` mirrors what
`TrajectoryTreeOptimizer.init_warm_start_cost_tree` / `init_cost_tree` (planners/mind/trajectory_tree.py:20-124) compute
per trajectory-tree node -- the grid frame of `gen_dist_field` (planners/ilqr/utils.py:5-22), the squared lane-distance
term, the exo-agent and ego covariance terms -- through ONE call of `mind_cost_fields` (csrc/cost_field.cu) for the whole
tree; the host keeps only the tree walk and the small per-node tables.  Returns numpy arrays in the layout the
reference's `PotentialField(offset, res, xx, yy, cost_field)` takes (binding shown in INTEGRATION.md).
No CPU fallback: needs the CUDA library and a CUDA device.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib


def grid_frame(ego_pos, grid_size, res):
    """utils.py:7-14: offset of cell (0,0), cell-centre coordinates x [gx], y [gy] (numpy forms them as the reference does)"""
    gx, gy = int(grid_size[0]), int(grid_size[1])
    size = ((gx - 1) * res, (gy - 1) * res)
    off = np.array([ego_pos[0] - 0.5 * size[0], ego_pos[1] - 0.5 * size[1]])
    return off, np.linspace(0.0, size[0], gx) + off[0], np.linspace(0.0, size[1], gy) + off[1]


def walk(scen_tree):
    """trajectory_tree.py:31-52: creation order of the trajectory-tree nodes: scenario nodes from a LIFO stack, one node
    per EVEN step.  Yields (scenario node, step, index, parent index); the root state has index -1."""
    last_of, stack, count = {}, [scen_tree.get_root()], 0
    while stack:
        node = stack.pop()
        last = last_of[node.parent_key] if node.parent_key is not None else -1
        for i in range(0, node.data[1].shape[1], 2):
            yield node, i, count, last
            last = count
            count += 1
        last_of[node.key] = count - 1
        stack.extend(scen_tree.get_node(k) for k in node.children_keys)


def node_tables(scen_tree, cfg, warm):
    """host half: per trajectory-tree node the coefficient of d^2, the actor centres [n,Na,2] and radii [n,Na] (fp32 sums
    as numpy forms `covs[e, i] + offset`), parent links and probabilities, in creation order"""
    coef, means, radii, links, probs = [], [], [], [], []
    for node, i, idx, last in walk(scen_tree):
        prob, trajs, covs = node.data[0], node.data[1], node.data[2]
        coef.append(float(cfg["w_tgt"] * prob))
        probs.append(prob)
        links.append((idx, last))
        if not warm:
            means.append(np.asarray(trajs[:, i], dtype=np.float64))
            r = (covs[:, i, 0] + np.float32(cfg["w_exo_cov_offset"])).astype(np.float64)
            r[0] = np.float64(covs[0, i, 0] + np.float32(cfg["w_ego_cov_offset"]))
            radii.append(r)
    return (np.array(coef), np.stack(means) if means else None, np.stack(radii) if radii else None, links, probs)


def cost_fields(scen_tree, x0, target_lane, cfg, device, warm=False, stream=None):
    """scen_tree: planners.basic.tree.Tree with node.data = [prob, trajs (Na,dur,2) f32, covs (Na,dur,1) f32, tgt_pts];
    x0: initial state (position in x0[:2]); cfg: the optimiser's `w_opt_cfg` (warm) or `opt_cfg` dict.
    Returns dict(offset [2], xx, yy [gy,gx], fields [n,gy,gx] fp64 numpy, links [(index, parent index)], probs [n])."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("cost_fields needs a CUDA device (no CPU fallback)")
    L = _lib.load()
    off, xs, ys = grid_frame(x0, cfg["smooth_grid_size"], cfg["smooth_grid_res"])
    gx, gy = len(xs), len(ys)
    coef, means, radii, links, probs = node_tables(scen_tree, cfg, warm)
    n, na = len(coef), (0 if means is None else means.shape[1])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
    d_xs, d_ys, d_lane, d_coef = dev(xs), dev(ys), dev(np.asarray(target_lane)[:, :2]), dev(coef)
    d_mean = dev(means) if na else None
    d_rad = dev(radii) if na else None
    quad = torch.empty(gy, gx, dtype=torch.float64, device=device)
    fields = torch.empty(n, gy, gx, dtype=torch.float64, device=device)
    a = _lib.MindCostFields()
    a.gx, a.gy, a.xs, a.ys = gx, gy, d_xs.data_ptr(), d_ys.data_ptr()
    a.n_lane_pts, a.lane = d_lane.shape[0], d_lane.data_ptr()
    a.n_nodes, a.n_actor = n, na
    a.coef_tgt = d_coef.data_ptr()
    a.mean, a.radius = (d_mean.data_ptr(), d_rad.data_ptr()) if na else (None, None)
    a.w_ego, a.w_exo = float(cfg.get("w_ego", 0.0)), float(cfg.get("w_exo", 0.0))
    a.exo_cost_offset = float(cfg.get("w_exo_cost_offset", 0.0))
    a.quad, a.fields = quad.data_ptr(), fields.data_ptr()
    st = stream if stream is not None else torch.cuda.current_stream(device).cuda_stream
    if L.mind_cost_fields(C.byref(a), C.c_void_p(st)) != 0:
        raise RuntimeError(L.mind_cost_fields_last_error().decode())
    xx, yy = np.meshgrid(xs, ys)
    return dict(offset=off, xx=xx, yy=yy, fields=fields.cpu().numpy(), quad=quad.cpu().numpy(), links=links, probs=probs)
