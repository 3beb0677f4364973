"""The slice of `av2==0.2.1` MIND calls, restated with numpy / pandas:

* `av2.map.map_api.ArgoverseStaticMap.from_json(path)` -> `.vector_lane_segments` (dict id -> LaneSegment, JSON order),
  `.get_lane_segment_centerline(id)` = midpoint of the two lane boundaries after each has been resampled to 10 points
  equally spaced in (3-D chordal) arc length  [av2.geometry.interpolate.interp_arc / compute_midpoint_line];
* `av2.map.lane_segment.{LaneType, LaneMarkType, LaneSegment}`;
* `av2.datasets.motion_forecasting.data_schema.{ObjectType, TrackCategory, ObjectState, Track, ArgoverseScenario}`;
* `av2.datasets.motion_forecasting.scenario_serialization.load_argoverse_scenario_parquet(path)`.

Used by: common/semantic_map.py:18,63; planners/mind/utils.py:298-311,345-483; loader.py:70-90; planner.py:58-90;
agent.py:86-98.  Parity against the real package is unpinned (it is not in the image).
"""
import enum
import json
import types
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

NUM_CENTERLINE_INTERP_PTS = 10


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
def interp_arc(t: int, points: np.ndarray) -> np.ndarray:
    """t points equally spaced in chordal arc length along the polyline `points` [n, d] (linear inside a segment)."""
    points = np.asarray(points, dtype=np.float64)
    if points.ndim != 2:
        raise ValueError("Input array must be (N,2) or (N,3) in shape.")
    n = points.shape[0]
    eq = np.linspace(0, 1, t)
    chord = np.linalg.norm(np.diff(points, axis=0), axis=1)
    chord = chord / np.sum(chord)
    cum = np.zeros(len(chord) + 1)
    cum[1:] = np.cumsum(chord)
    bins = np.digitize(eq, bins=cum).astype(int)
    bins[np.where((bins <= 0) | (eq <= 0))] = 1
    bins[np.where((bins >= n) | (eq >= 1))] = n - 1
    s = np.divide(eq - cum[bins - 1], chord[bins - 1])
    return points[bins - 1, :] + (points[bins, :] - points[bins - 1, :]) * s.reshape(-1, 1)


def compute_midpoint_line(left: np.ndarray, right: np.ndarray, num_interp_pts: int = NUM_CENTERLINE_INTERP_PTS):
    if left.ndim != 2 or right.ndim != 2:
        raise ValueError("Each polyline must be (N,2) or (N,3)")
    le, re = interp_arc(num_interp_pts, left), interp_arc(num_interp_pts, right)
    return (le + re) / 2.0, float(np.linalg.norm(le - re, axis=1).mean())


# ---------------------------------------------------------------------------------------------
# av2.map.lane_segment
# ---------------------------------------------------------------------------------------------
class LaneType(str, enum.Enum):
    VEHICLE = "VEHICLE"
    BIKE = "BIKE"
    BUS = "BUS"


class LaneMarkType(str, enum.Enum):
    DASH_SOLID_YELLOW = "DASH_SOLID_YELLOW"
    DASH_SOLID_WHITE = "DASH_SOLID_WHITE"
    DASHED_WHITE = "DASHED_WHITE"
    DASHED_YELLOW = "DASHED_YELLOW"
    DOUBLE_SOLID_YELLOW = "DOUBLE_SOLID_YELLOW"
    DOUBLE_SOLID_WHITE = "DOUBLE_SOLID_WHITE"
    DOUBLE_DASH_YELLOW = "DOUBLE_DASH_YELLOW"
    DOUBLE_DASH_WHITE = "DOUBLE_DASH_WHITE"
    SOLID_YELLOW = "SOLID_YELLOW"
    SOLID_WHITE = "SOLID_WHITE"
    SOLID_DASH_WHITE = "SOLID_DASH_WHITE"
    SOLID_DASH_YELLOW = "SOLID_DASH_YELLOW"
    SOLID_BLUE = "SOLID_BLUE"
    NONE = "NONE"
    UNKNOWN = "UNKNOWN"


@dataclass
class Polyline:
    xyz: np.ndarray                                     # [n, 3]

    @classmethod
    def from_json_data(cls, pts):
        return cls(np.array([[p["x"], p["y"], p["z"]] for p in pts], dtype=np.float64))


@dataclass
class LaneSegment:
    id: int
    is_intersection: bool
    lane_type: LaneType
    right_lane_boundary: Polyline
    left_lane_boundary: Polyline
    right_mark_type: LaneMarkType
    left_mark_type: LaneMarkType
    predecessors: List[int]
    successors: List[int]
    right_neighbor_id: Optional[int] = None
    left_neighbor_id: Optional[int] = None

    @classmethod
    def from_dict(cls, d):
        return cls(id=d["id"], is_intersection=d["is_intersection"], lane_type=LaneType(d["lane_type"]),
                   right_lane_boundary=Polyline.from_json_data(d["right_lane_boundary"]),
                   left_lane_boundary=Polyline.from_json_data(d["left_lane_boundary"]),
                   right_mark_type=LaneMarkType(d["right_lane_mark_type"]),
                   left_mark_type=LaneMarkType(d["left_lane_mark_type"]),
                   right_neighbor_id=d["right_neighbor_id"], left_neighbor_id=d["left_neighbor_id"],
                   predecessors=d["predecessors"], successors=d["successors"])

    @property
    def polygon_boundary(self) -> np.ndarray:
        return np.vstack([self.right_lane_boundary.xyz, self.left_lane_boundary.xyz[::-1], self.right_lane_boundary.xyz[:1]])


# ---------------------------------------------------------------------------------------------
# av2.map.map_api
# ---------------------------------------------------------------------------------------------
class ArgoverseStaticMap:
    def __init__(self, log_id, vector_lane_segments, vector_drivable_areas=None, vector_pedestrian_crossings=None):
        self.log_id = log_id
        self.vector_lane_segments: Dict[int, LaneSegment] = vector_lane_segments
        self.vector_drivable_areas = vector_drivable_areas or {}
        self.vector_pedestrian_crossings = vector_pedestrian_crossings or {}

    @classmethod
    def from_json(cls, static_map_path):
        static_map_path = Path(static_map_path)
        log_id = static_map_path.stem.split("log_map_archive_")[-1]
        with open(static_map_path) as f:
            data = json.load(f)
        lanes = {ls["id"]: LaneSegment.from_dict(ls) for ls in data["lane_segments"].values()}
        return cls(log_id, lanes, data.get("drivable_areas"), data.get("pedestrian_crossings"))

    def get_lane_segment_centerline(self, lane_segment_id: int) -> np.ndarray:
        ls = self.vector_lane_segments[lane_segment_id]
        center, _ = compute_midpoint_line(ls.left_lane_boundary.xyz, ls.right_lane_boundary.xyz, NUM_CENTERLINE_INTERP_PTS)
        return center

    def get_scenario_lane_segments(self) -> List[LaneSegment]:
        return list(self.vector_lane_segments.values())

    def get_scenario_lane_segment_ids(self) -> List[int]:
        return list(self.vector_lane_segments.keys())


# ---------------------------------------------------------------------------------------------
# av2.datasets.motion_forecasting.data_schema
# ---------------------------------------------------------------------------------------------
class TrackCategory(enum.Enum):
    TRACK_FRAGMENT = 0
    UNSCORED_TRACK = 1
    SCORED_TRACK = 2
    FOCAL_TRACK = 3


class ObjectType(str, enum.Enum):
    VEHICLE = "vehicle"
    PEDESTRIAN = "pedestrian"
    MOTORCYCLIST = "motorcyclist"
    CYCLIST = "cyclist"
    BUS = "bus"
    STATIC = "static"
    BACKGROUND = "background"
    CONSTRUCTION = "construction"
    RIDERLESS_BICYCLE = "riderless_bicycle"
    UNKNOWN = "unknown"


@dataclass
class ObjectState:
    observed: bool
    timestep: int
    position: Tuple[float, float]
    heading: float
    velocity: Tuple[float, float]


@dataclass
class Track:
    track_id: str
    object_states: List[ObjectState]
    object_type: ObjectType
    category: TrackCategory


@dataclass
class ArgoverseScenario:
    scenario_id: str
    timestamps_ns: np.ndarray
    tracks: List[Track]
    focal_track_id: str
    city_name: str
    map_id: Optional[int] = None
    slice_id: Optional[str] = None


# ---------------------------------------------------------------------------------------------
# av2.datasets.motion_forecasting.scenario_serialization
# ---------------------------------------------------------------------------------------------
def _tracks_from_table(df) -> List[Track]:
    tracks = []
    for track_id, tdf in df.groupby("track_id"):
        states = [ObjectState(observed=bool(o), timestep=int(t), position=(float(px), float(py)), heading=float(h),
                              velocity=(float(vx), float(vy)))
                  for o, t, px, py, h, vx, vy in zip(tdf["observed"].values, tdf["timestep"].values, tdf["position_x"].values,
                                                     tdf["position_y"].values, tdf["heading"].values, tdf["velocity_x"].values,
                                                     tdf["velocity_y"].values)]
        tracks.append(Track(track_id=str(track_id), object_states=states, object_type=ObjectType(tdf["object_type"].iloc[0]),
                            category=TrackCategory(int(tdf["object_category"].iloc[0]))))
    return tracks


def load_argoverse_scenario_parquet(scenario_path) -> ArgoverseScenario:
    import pandas as pd
    scenario_path = Path(scenario_path)
    if not scenario_path.exists():
        raise FileNotFoundError("No scenario exists at location: %s." % scenario_path)
    df = pd.read_parquet(scenario_path)
    ts = np.linspace(df["start_timestamp"].iloc[0], df["end_timestamp"].iloc[0], num=int(df["num_timestamps"].iloc[0]))
    return ArgoverseScenario(scenario_id=str(df["scenario_id"].iloc[0]), timestamps_ns=ts, tracks=_tracks_from_table(df),
                             focal_track_id=str(df["focal_track_id"].iloc[0]), city_name=str(df["city"].iloc[0]))


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


lane_segment = _module("av2.map.lane_segment", LaneType=LaneType, LaneMarkType=LaneMarkType, LaneSegment=LaneSegment,
                       Polyline=Polyline)
map_api = _module("av2.map.map_api", ArgoverseStaticMap=ArgoverseStaticMap, LaneSegment=LaneSegment)
data_schema = _module("av2.datasets.motion_forecasting.data_schema", ObjectType=ObjectType, TrackCategory=TrackCategory,
                      ObjectState=ObjectState, Track=Track, ArgoverseScenario=ArgoverseScenario)
scenario_serialization = _module("av2.datasets.motion_forecasting.scenario_serialization",
                                 load_argoverse_scenario_parquet=load_argoverse_scenario_parquet)
