"""CPU restatement (numpy fp64) of the cost fields MIND's trajectory-tree optimiser builds for every node of the
trajectory tree.  Test infrastructure: only tests/ may import it.

Follows  planners/ilqr/utils.py:5-22 (gen_dist_field: grid frame centred on the ego position, distance of every cell
centre to the target-lane polyline = min over segments of the clamped point-segment distance, common/geometry.py:70-78),
planners/mind/trajectory_tree.py:20-56 (warm start: w_tgt * prob * d^2) and :58-124 (w_tgt * prob * d^2
+ w_exo * sum_exo [max(r_exo - |p - mu_exo|, 0) (+ cost offset where positive)] + w_ego * max(|p - mu_ego| - r_ego, 0),
one field per EVEN step of every scenario-tree node, scenario nodes visited depth first from a LIFO stack).
Pinned against the reference's own functions in tests/test_cost_field_cpu.py (live, build container) and against
tests/golden/cost_fields_demo_2.npz dumped from them (oracle/make_golden_cost_fields.py).
"""
import numpy as np


def grid_frame(ego_pos, grid_size, res):
    """utils.py:7-14: offset of cell (0, 0), xx / yy [gy, gx] (xx varies along columns)"""
    gx, gy = int(grid_size[0]), int(grid_size[1])
    size = ((gx - 1) * res, (gy - 1) * res)
    off = np.array([ego_pos[0] - 0.5 * size[0], ego_pos[1] - 0.5 * size[1]])
    x = np.linspace(0.0, size[0], gx) + off[0]
    y = np.linspace(0.0, size[1], gy) + off[1]
    xx, yy = np.meshgrid(x, y)
    return off, xx, yy


def lane_distance(xx, yy, polyline):
    """utils.py:16-22 + geometry.py:70-78, all segments at once per cell"""
    p = np.asarray(polyline, dtype=np.float64)
    best = np.full(xx.shape, np.inf)
    for a, b in zip(p[:-1], p[1:]):
        lx, ly = b[0] - a[0], b[1] - a[1]
        t = np.clip(((xx - a[0]) * lx + (yy - a[1]) * ly) / (lx * lx + ly * ly), 0.0, 1.0)
        dx, dy = xx - (a[0] + t * lx), yy - (a[1] + t * ly)
        best = np.minimum(best, np.sqrt(dx * dx + dy * dy))
    return best


def walk(nodes, root_key):
    """trajectory_tree.py:31-52 / :71-121: order in which trajectory-tree nodes are created.
    nodes: {key: (parent_key, prob, trajs [Na,dur,2], covs [Na,dur,1], children_keys)}.
    Yields (scenario key, step i, prob, index, parent index); indices count from 0, the root state is -1."""
    last_of, stack, count = {}, [root_key], 0
    while stack:
        key = stack.pop()
        parent, prob, trajs, covs, children = nodes[key]
        last = last_of[parent] if parent is not None else -1
        for i in range(trajs.shape[1]):
            if i % 2 == 1:
                continue
            yield key, i, prob, count, last
            last = count
            count += 1
        last_of[key] = count - 1
        stack.extend(children)


def node_inputs(nodes, root_key, cfg, warm):
    """per trajectory-tree node: coefficient of d^2, actor centres [Na,2] and radii [Na] (fp32 sums as numpy forms them:
    covs is an fp32 array, the offsets are Python floats), in creation order"""
    coef, mean, rad, links = [], [], [], []
    for key, i, prob, idx, last in walk(nodes, root_key):
        _, _, trajs, covs, _ = nodes[key]
        coef.append(float(cfg["w_tgt"] * prob))
        links.append((idx, last))
        if not warm:
            mean.append(np.asarray(trajs[:, i], dtype=np.float64))
            r = np.empty(trajs.shape[0])
            r[0] = (covs[0, i] + cfg["w_ego_cov_offset"])[0]
            for e in range(1, trajs.shape[0]):
                r[e] = (covs[e, i] + cfg["w_exo_cov_offset"])[0]
            rad.append(r)
    return np.array(coef), (np.stack(mean) if mean else None), (np.stack(rad) if rad else None), links


def cost_fields(nodes, root_key, x0, target_lane, cfg, warm=False):
    """Returns (offset, xx, yy, fields [n, gy, gx], links [(index, parent index)])."""
    off, xx, yy = grid_frame(x0, cfg["smooth_grid_size"], cfg["smooth_grid_res"])
    quad = lane_distance(xx, yy, target_lane) ** 2
    coef, mean, rad, links = node_inputs(nodes, root_key, cfg, warm)
    out = np.empty((len(coef),) + xx.shape)
    for n in range(len(coef)):
        if warm:
            out[n] = coef[n] * quad
            continue
        d = np.sqrt((xx - mean[n, 0, 0]) ** 2 + (yy - mean[n, 0, 1]) ** 2)
        ego = np.maximum(d - rad[n, 0], 0.0)
        acc = np.zeros_like(xx)
        for e in range(1, mean.shape[1]):
            f = np.maximum(rad[n, e] - np.sqrt((xx - mean[n, e, 0]) ** 2 + (yy - mean[n, e, 1]) ** 2), 0.0)
            acc += np.where(f > 0, f + cfg["w_exo_cost_offset"], f)
        out[n] = coef[n] * quad + cfg["w_exo"] * acc + cfg["w_ego"] * ego
    return off, xx, yy, out, links
