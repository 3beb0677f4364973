"""Minimal stand-ins for the third-party packages MIND imports around the hot path.

The reference needs `av2==0.2.1` (map JSON -> lane segments / 10-point centrelines, scenario parquet -> tracks),
`shapely==2.0.6` (arc-length resampling of centrelines) and `Theano==1.0.5` (symbolic bicycle model -> f, f_x, f_u of the
tree iLQR).  None of them
is a hot-path dependency, but without them `ScenarioTreeGenerator.process_data`, `SemanticMap` and `ArgoAgentLoader`
cannot be imported, so real Argoverse-2 scenes cannot reach the predictor.  `install()` registers pure numpy / pandas
implementations of exactly the API surface MIND touches under the original module names, unless the real packages are
importable.  Parity against the real packages is UNPINNED (they are not in this image; the algorithms are restated from
their published sources, see each module's docstring).
"""
import importlib
import importlib.util
import sys
import types


def _have(name: str) -> bool:
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


_installed = None      # decision of the first install(): later calls return it instead of re-probing / re-registering


def install(force: bool = False) -> dict:
    """Register the stand-ins in sys.modules (once).  Returns {package: 'real' | 'lite'}."""
    global _installed
    if _installed is not None and not force:
        return dict(_installed)
    used = {}
    if force or not _have("av2"):
        from . import av2_lite
        pk = types.ModuleType("av2"); pk.__path__ = []
        mp = types.ModuleType("av2.map"); mp.__path__ = []
        ds = types.ModuleType("av2.datasets"); ds.__path__ = []
        mf = types.ModuleType("av2.datasets.motion_forecasting"); mf.__path__ = []
        mods = {"av2": pk, "av2.map": mp, "av2.map.map_api": av2_lite.map_api, "av2.map.lane_segment": av2_lite.lane_segment,
                "av2.datasets": ds, "av2.datasets.motion_forecasting": mf,
                "av2.datasets.motion_forecasting.data_schema": av2_lite.data_schema,
                "av2.datasets.motion_forecasting.scenario_serialization": av2_lite.scenario_serialization}
        pk.map, pk.datasets = mp, ds
        mp.map_api, mp.lane_segment = av2_lite.map_api, av2_lite.lane_segment
        ds.motion_forecasting = mf
        mf.data_schema, mf.scenario_serialization = av2_lite.data_schema, av2_lite.scenario_serialization
        sys.modules.update(mods)
        used["av2"] = "lite"
    else:
        used["av2"] = "real"
    if force or not _have("shapely"):
        from . import shapely_lite
        pk = types.ModuleType("shapely"); pk.__path__ = []
        pk.geometry = shapely_lite.geometry
        pk.ops = shapely_lite.ops
        sys.modules.update({"shapely": pk, "shapely.geometry": shapely_lite.geometry, "shapely.ops": shapely_lite.ops})
        used["shapely"] = "lite"
    else:
        used["shapely"] = "real"
    if force or not _have("theano"):
        from . import theano_lite
        pk = types.ModuleType("theano"); pk.__path__ = []
        pk.tensor, pk.function = theano_lite.tensor, theano_lite.function
        sys.modules.update({"theano": pk, "theano.tensor": theano_lite.tensor})
        used["theano"] = "lite"
    else:
        used["theano"] = "real"
    if force or not _have("matplotlib"):
        from . import mpl_stub
        sys.modules.update(mpl_stub.modules())
        used["matplotlib"] = "stub (no rendering)"
    else:
        used["matplotlib"] = "real"
    _installed = dict(used)
    return used
