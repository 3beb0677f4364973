"""`shapely.geometry.LineString` / `Point` as far as MIND uses them (planners/mind/utils.py:357-369:
`LineString(pts).length`, `.interpolate(s)`, `.coords`, and a LineString built from interpolated Points).
Restates GEOS' linear referencing: walk the segments, interpolate linearly inside the segment that contains the
distance, clamp to the end points."""
import types

import numpy as np


class Point:
    def __init__(self, x, y=None):
        if y is None:
            x, y = x[0], x[1]
        self.x, self.y = float(x), float(y)

    @property
    def coords(self):
        return [(self.x, self.y)]

    def __iter__(self):
        return iter((self.x, self.y))

    def __repr__(self):
        return "POINT (%r %r)" % (self.x, self.y)


class LineString:
    def __init__(self, pts):
        rows = []
        for p in pts:
            if isinstance(p, Point):
                rows.append((p.x, p.y))
            else:
                rows.append((float(p[0]), float(p[1])))
        self._xy = np.asarray(rows, dtype=np.float64).reshape(-1, 2)
        if len(self._xy) == 1:
            raise ValueError("LineStrings must have at least 2 coordinate tuples")
        self._seg = np.linalg.norm(np.diff(self._xy, axis=0), axis=1) if len(self._xy) else np.zeros(0)
        self._cum = np.concatenate([[0.0], np.cumsum(self._seg)])

    @property
    def length(self) -> float:
        return float(self._cum[-1])

    @property
    def coords(self):
        return [tuple(r) for r in self._xy]

    def interpolate(self, distance, normalized=False) -> Point:
        d = float(distance) * (self.length if normalized else 1.0)
        if d < 0:
            d = max(0.0, self.length + d)          # shapely: negative distances are measured from the end
        if d <= 0.0:
            return Point(self._xy[0])
        if d >= self._cum[-1]:
            return Point(self._xy[-1])
        k = int(np.searchsorted(self._cum, d, side="right")) - 1
        k = min(max(k, 0), len(self._seg) - 1)
        t = (d - self._cum[k]) / self._seg[k] if self._seg[k] > 0 else 0.0
        return Point(self._xy[k] + t * (self._xy[k + 1] - self._xy[k]))


class Polygon:          # imported by common/visualization.py only; never constructed on the data path
    def __init__(self, *a, **k):
        raise NotImplementedError("shapely_lite: Polygon is rendering-only")


def _unary_union(*a, **k):
    raise NotImplementedError("shapely_lite: unary_union is rendering-only")


geometry = types.ModuleType("shapely.geometry")
geometry.LineString, geometry.Point, geometry.Polygon = LineString, Point, Polygon
ops = types.ModuleType("shapely.ops")
ops.unary_union = _unary_union
