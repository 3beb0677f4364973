"""Import the UNMODIFIED reference hot-path modules from /root/reference.

Test infrastructure only.  Works only in the build container (the GPU box has
no /root/reference); everything that must travel is dumped to tests/golden/
by oracle/make_golden.py.

The reference imports shapely / av2 symbols at module import time
(planners/mind/utils.py:5-7) although no hot-path function touches them; we
register empty stand-in modules so the import succeeds.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MIND_REFERENCE_ROOT", "/root/reference")
CKPT = os.path.join(REF_ROOT, "planners/mind/check_points/20240121-172745.tar")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "planners", "mind"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_shims():
    _stub("shapely")
    _stub("shapely.geometry", LineString=object)
    _stub("av2")
    _stub("av2.map")
    _stub("av2.map.lane_segment", LaneType=object, LaneMarkType=object)
    _stub("av2.datasets")
    _stub("av2.datasets.motion_forecasting")
    _stub("av2.datasets.motion_forecasting.data_schema", ObjectType=object)


def load():
    """Returns a namespace with the reference classes/functions of the path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.network = importlib.import_module("planners.mind.networks.network")
    ns.utils = importlib.import_module("planners.mind.utils")
    ns.scenario_tree = importlib.import_module("planners.mind.scenario_tree")
    ns.tree = importlib.import_module("planners.basic.tree")
    ns.net_cfg = importlib.import_module("planners.mind.configs.networks.net_cfg")
    ns.plan_cfg = importlib.import_module("planners.mind.configs.planning.demo_2")
    return ns


def build_reference_net(state_dict=None, device="cpu"):
    """ScenePredNet as planners/mind/planner.py:42-49 builds it."""
    import torch
    ns = load()
    cfg = ns.net_cfg.NetCfg().get_net_cfg()
    net = ns.network.ScenePredNet(cfg, torch.device(device))
    if state_dict is None:
        state_dict = torch.load(CKPT, map_location="cpu")["state_dict"]
    net.load_state_dict(state_dict)
    net = net.to(torch.device(device))
    net.eval()
    return net, cfg
