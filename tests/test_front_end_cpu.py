"""CPU: the compat layer (av2 / shapely stand-ins) and the product's scene front end.

* unit tests of the stand-ins (arc-length interpolation, LineString referencing, parquet / map loaders on a synthetic
  scene written to tmp_path) run everywhere;
* the front end vs the UNMODIFIED reference's process_data on the four demo scenes needs /root/reference (build
  container only) and is skipped elsewhere: the comparison is bit-exact on every tensor of the collated dict."""
import json
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ref_loader


def test_interp_arc_and_midpoint():
    from mind_b200.compat import av2_lite as A
    pts = np.array([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0], [3.0, 4.0, 0.0]])            # length 7
    out = A.interp_arc(8, pts)
    assert out.shape == (8, 3) and np.allclose(out[0], pts[0]) and np.allclose(out[-1], pts[-1])
    seg = np.linalg.norm(np.diff(out, axis=0), axis=1)
    assert np.allclose(out[3], [3.0, 0.0, 0.0]) and np.allclose(seg[:3], 1.0) and np.allclose(seg[3:], 1.0)
    mid, width = A.compute_midpoint_line(pts, pts + np.array([0.0, 0.0, 2.0]), 10)
    assert mid.shape == (10, 3) and np.allclose(mid[:, 2], 1.0) and abs(width - 2.0) < 1e-12


def test_linestring_referencing():
    from mind_b200.compat.shapely_lite import LineString, Point
    ls = LineString([(0, 0), (3, 0), (3, 4)])
    assert ls.length == 7.0
    assert tuple(ls.interpolate(0.0)) == (0.0, 0.0) and tuple(ls.interpolate(7.0)) == (3.0, 4.0) and tuple(ls.interpolate(99)) == (3.0, 4.0)
    assert tuple(ls.interpolate(3.0)) == (3.0, 0.0) and tuple(ls.interpolate(5.0)) == (3.0, 2.0)
    assert tuple(ls.interpolate(0.5, normalized=True)) == (3.0, 0.5)
    ls2 = LineString([ls.interpolate(s) for s in (0.0, 1.5, 3.0)])
    assert ls2.coords == [(0.0, 0.0), (1.5, 0.0), (3.0, 0.0)] and isinstance(ls.interpolate(1), Point)
    with pytest.raises(ValueError):
        LineString([(0, 0)])


def test_loaders_on_a_synthetic_log(tmp_path):
    import pandas as pd
    from mind_b200.compat import av2_lite as A
    pt = lambda x, y: {"x": x, "y": y, "z": 0.0}
    lane = lambda i, y, pred, succ: {"id": i, "is_intersection": False, "lane_type": "VEHICLE", "left_lane_mark_type": "DASHED_WHITE",
                                     "right_lane_mark_type": "SOLID_WHITE", "left_neighbor_id": None, "right_neighbor_id": 7,
                                     "predecessors": pred, "successors": succ,
                                     "left_lane_boundary": [pt(0, y + 2), pt(30, y + 2)], "right_lane_boundary": [pt(0, y - 2), pt(10, y - 2), pt(30, y - 2)]}
    mp = tmp_path / "log_map_archive_x.json"
    mp.write_text(json.dumps({"lane_segments": {"5": lane(5, 0.0, [], [6]), "6": lane(6, 10.0, [5], [])}, "drivable_areas": {}, "pedestrian_crossings": {}}))
    m = A.ArgoverseStaticMap.from_json(mp)
    assert list(m.vector_lane_segments) == [5, 6] and m.vector_lane_segments[5].right_mark_type == A.LaneMarkType.SOLID_WHITE
    cl = m.get_lane_segment_centerline(6)
    assert cl.shape == (10, 3) and np.allclose(cl[:, 1], 10.0) and np.allclose(cl[:, 0], np.linspace(0, 30, 10))
    rows = []
    for tid, cat, typ in (("AV", 0, "vehicle"), ("12", 3, "pedestrian")):
        for t in range(3):
            rows.append(dict(observed=t < 2, track_id=tid, object_type=typ, object_category=cat, timestep=t, position_x=1.0 * t, position_y=2.0,
                             heading=0.1, velocity_x=1.0, velocity_y=0.0, scenario_id="x", start_timestamp=0.0, end_timestamp=2.0,
                             num_timestamps=3, focal_track_id="12", city="nowhere"))
    pq = tmp_path / "scenario_x.parquet"
    pd.DataFrame(rows).to_parquet(pq)
    sc = A.load_argoverse_scenario_parquet(pq)
    assert sc.focal_track_id == "12" and [t.track_id for t in sc.tracks] == ["12", "AV"] and len(sc.timestamps_ns) == 3
    assert sc.tracks[0].category == A.TrackCategory.FOCAL_TRACK and sc.tracks[0].object_type == A.ObjectType.PEDESTRIAN
    assert sc.tracks[1].object_states[2].observed is False and sc.tracks[1].object_states[1].position == (1.0, 2.0)


def test_agent_trajectories_padding_and_order():
    from mind_b200.compat import av2_lite as A
    from mind_b200.front_end import agent_trajectories
    st = lambda obs, t, x: A.ObjectState(obs, t, (x, 0.0), 0.5, (1.0, 0.0))
    obs = {"7": A.Track("7", [st(False, 0, 0.0), st(True, 1, 1.0), st(False, 2, 1.0), st(True, 3, 3.0)], A.ObjectType.CYCLIST, A.TrackCategory.TRACK_FRAGMENT),
           "gone": A.Track("gone", [st(True, 0, 5.0), st(False, 1, 5.0)], A.ObjectType.VEHICLE, A.TrackCategory.TRACK_FRAGMENT),
           "AV": A.Track("AV", [st(True, t, float(t)) for t in range(50)], A.ObjectType.VEHICLE, A.TrackCategory.FOCAL_TRACK)}
    tr = agent_trajectories(obs)
    assert tr["tid"] == ["AV", "7"] and tr["cat"] == ["av", "exo"]
    assert tr["pos"].shape == (2, 50, 2) and tr["type"].dtype == torch.int16 and tr["flags"].dtype == torch.int16
    assert tr["flags"][1].tolist() == [0] * 46 + [0, 1, 0, 1]
    assert tr["pos"][1, :, 0].tolist() == [1.0] * 48 + [1.0, 3.0]            # leading steps <- first seen, gap <- previous seen
    assert tr["vel"][1, :, 0].tolist() == [0.0] * 47 + [1.0, 0.0, 1.0]
    assert tr["type"][1, 47].tolist() == [0, 0, 0, 1, 0, 0, 0] and tr["type"][1, 48].tolist() == [0] * 7


def _same(a, b, path="data"):
    if torch.is_tensor(a):
        assert torch.is_tensor(b) and a.dtype == b.dtype and a.shape == b.shape, path
        assert torch.equal(a, b), (path, (a.float() - b.float()).abs().max().item())
    elif isinstance(a, dict):
        assert set(a) == set(b), (path, set(a) ^ set(b))
        for k in a:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, "%s[%d]" % (path, i))
    else:
        assert a == b, path


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("name", ["demo_1", "demo_2", "demo_3", "demo_4"])
def test_front_end_vs_reference_process_data(name):
    """live reference loader / agents / planner observation -> (a) its process_data equals the committed golden dict,
    (b) the product front end returns the same dict bit for bit; the lane-graph cache is exercised by a second call."""
    from oracle import make_golden_real as M
    from mind_b200.front_end import ArgoFrontEnd
    from mind_b200 import plumbing as P
    gold = torch.load(os.path.join(GOLDEN, "real_%s.pt" % name), weights_only=False)
    ns = M.load_reference_sim()
    cfg = json.load(open(os.path.join(ref_loader.REF_ROOT, "configs", name + ".json")))
    ego, _ = M.run_until(ns, cfg, gold["t_plan"])
    pl = ego.planner
    lane, info = pl.resample_target_lane(ego.lcl_smp)
    assert np.array_equal(np.asarray(lane), gold["lane"])
    gen = types.SimpleNamespace(device=torch.device("cpu"), lane_graph=None, config=types.SimpleNamespace(tar_time_ahead=5.0),
                                target_lane=torch.from_numpy(np.array(lane)), target_lane_info=P.pack_target_lane_info(info))
    fe = ArgoFrontEnd(gen)
    for _ in range(2):
        got = fe(ego.lcl_smp, pl.agent_obs)
        _same(got, gold["data"])
        _same(gen.lane_graph, gold["graph"])
    assert len(fe._maps) == 1


def test_theano_lite_bicycle_model_jacobians():
    """the symbolic slice MIND's tree iLQR needs (trajectory_tree.py:153-177 through ilqr/autodiff.py): exact Jacobians of
    the 6-state kinematic bicycle model, checked against central finite differences; numpy scalars mix into expressions"""
    from mind_b200.compat import theano_lite as TL
    T = TL.tensor
    dt, wb = np.float64(0.2), 2.5
    x = [T.dscalar(n) for n in ("x", "y", "v", "q", "a", "theta")]
    u = [T.dscalar(n) for n in ("da", "dtheta")]
    i = T.dscalar("i")
    f = T.stack([x[0] + x[2] * T.cos(x[3]) * dt, x[1] + x[2] * T.sin(x[3]) * dt, x[2] + x[4] * dt,
                 x[3] + x[2] / wb * T.tan(x[5]) * dt, x[4] + u[0] * dt, x[5] + u[1] * dt])
    inputs = np.hstack([x, u, i]).tolist()                                   # dynamics.py:167-168
    wrt = np.hstack([x, u]).tolist()
    J = T.stack([T.grad(f[k], wrt, disconnected_inputs="ignore") for k in range(6)])     # autodiff.jacobian_vector
    F = TL.function(inputs, f, on_unused_input="ignore", name="f")
    Fx = TL.function(inputs, J[:, :6], on_unused_input="ignore", name="f_x")
    Fu = TL.function(inputs, J[:, 6:], on_unused_input="ignore", name="f_u")
    rng = np.random.default_rng(0)
    for _ in range(5):
        z = np.concatenate([rng.normal(size=3) * 5, rng.uniform(-1, 1, 3), rng.normal(size=2), [0.0]])
        fz = F(*z)
        assert fz.shape == (6,) and abs(fz[0] - (z[0] + z[2] * np.cos(z[3]) * 0.2)) < 1e-12
        num = np.zeros((6, 8))
        for k in range(8):
            e = np.zeros(9); e[k] = 1e-6
            num[:, k] = (F(*(z + e)) - F(*(z - e))) / 2e-6
        assert Fx(*z).shape == (6, 6) and Fu(*z).shape == (6, 2)
        assert np.abs(Fx(*z) - num[:, :6]).max() < 1e-6 and np.abs(Fu(*z) - num[:, 6:]).max() < 1e-6
    with pytest.raises(NotImplementedError):
        T.dvector("v")
    with pytest.raises(NotImplementedError):
        T.grad(cost=None, wrt=x, known_grads={})


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree (build container only)")
def test_reference_trajectory_optimizer_builds_on_theano_lite():
    """the UNMODIFIED TrajectoryTreeOptimizer constructs its AutoDiffDynamics on the stand-in and reproduces one Euler step"""
    import sys
    from mind_b200 import compat
    assert compat.install()["theano"] in ("lite", "real")
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    from planners.mind.trajectory_tree import TrajectoryTreeOptimizer
    from planners.mind.configs.planning.demo_1 import TrajTreeCfg
    cfg = TrajTreeCfg()
    dyn = TrajectoryTreeOptimizer(cfg).ilqr.dynamics
    x, u = np.array([1.0, 2.0, 5.0, 0.3, 0.5, 0.1]), np.array([0.2, -0.1])
    nxt = dyn.f(x, u, 0)
    want = np.array([x[0] + x[2] * np.cos(x[3]) * cfg.dt, x[1] + x[2] * np.sin(x[3]) * cfg.dt, x[2] + x[4] * cfg.dt,
                     x[3] + x[2] / 2.5 * np.tan(x[5]) * cfg.dt, x[4] + u[0] * cfg.dt, x[5] + u[1] * cfg.dt])
    assert np.abs(nxt - want).max() < 1e-12 and dyn.f_x(x, u, 0).shape == (6, 6) and dyn.f_u(x, u, 0).shape == (6, 2)


@pytest.mark.skipif(not ref_loader.available(), reason="needs the demo maps of the reference tree (build container only)")
def test_centerlines_agree_with_the_maps_own_centerline_field():
    """weak pin of the unpinned av2 boundary: every AV2 map JSON also ships a `centerline` polyline per lane segment; the
    10-point midpoint line of the resampled boundaries (what av2's get_lane_segment_centerline returns) must run along it"""
    import glob
    from mind_b200.compat import av2_lite as A
    worst = 0.0
    n = 0
    for path in glob.glob(os.path.join(ref_loader.REF_ROOT, "data", "*", "log_map_archive_*.json")):
        raw = json.load(open(path))["lane_segments"]
        m = A.ArgoverseStaticMap.from_json(path)
        for key, ls in raw.items():
            ref = np.array([[p["x"], p["y"]] for p in ls["centerline"]])
            got = m.get_lane_segment_centerline(ls["id"])[:, :2]
            assert got.shape == (10, 2)
            # distance of each of our points to the map's polyline
            a, b = ref[:-1], ref[1:]
            d = b - a
            t = np.clip(((got[:, None] - a[None]) * d[None]).sum(-1) / (d * d).sum(-1)[None], 0, 1)
            dist = np.linalg.norm(got[:, None] - (a[None] + t[..., None] * d[None]), axis=-1).min(axis=1)
            worst = max(worst, float(dist.max()))
            n += 1
    assert n > 100 and worst < 0.6, (n, worst)          # lanes are ~3.5 m wide; the two constructions differ by centimetres


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree (build container only)")
@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode of the plugin path")
def test_reference_plugin_mechanism_reaches_the_b200_class(tmp_path):
    """the drop-in is configuration only: a planner JSON whose `network_config` names mind_b200.integration.net_cfg_b200
    makes the UNMODIFIED MINDPlanner.init_network (planner.py:42-49) import ScenePredNetB200 -- which, without a CUDA
    device, refuses loudly instead of falling back"""
    import sys
    from mind_b200 import compat
    compat.install()
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    from mind_b200.integration.net_cfg_b200 import NetCfg
    cfg = NetCfg().get_net_cfg()
    assert cfg["network"] == "mind_b200.predictor:ScenePredNetB200" and cfg["n_scene_layer"] == 6 and cfg["g_num_modes"] == 6
    pc = json.load(open(os.path.join(ref_loader.REF_ROOT, "planners/mind/configs/demo_2.json")))
    pc["network_config"] = "mind_b200.integration.net_cfg_b200"
    path = tmp_path / "planner.json"
    path.write_text(json.dumps(pc))
    from planners.mind.planner import MINDPlanner
    cwd = os.getcwd()
    os.chdir(ref_loader.REF_ROOT)
    try:
        with pytest.raises(RuntimeError, match="CUDA"):
            MINDPlanner(str(path))
    finally:
        os.chdir(cwd)
