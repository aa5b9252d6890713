"""`Polygon.intersects` for convex polygons as a closed-set separating-axis test.

GEOS `intersects` is "not disjoint": polygons that only touch along an edge or at
a corner DO intersect, hence the strict `<` in the separation test below.
The reference only ever builds axis-aligned rectangles (merging_env.py:201-203).
"""


class Polygon:
    def __init__(self, shell):
        self.pts = [(float(x), float(y)) for x, y in shell]

    def _axes(self):
        n = len(self.pts)
        for i in range(n):
            x0, y0 = self.pts[i]
            x1, y1 = self.pts[(i + 1) % n]
            yield (-(y1 - y0), x1 - x0)

    @staticmethod
    def _project(pts, ax):
        d = [p[0] * ax[0] + p[1] * ax[1] for p in pts]
        return min(d), max(d)

    def intersects(self, other):
        for ax in list(self._axes()) + list(other._axes()):
            if ax == (0.0, 0.0):
                continue
            a0, a1 = self._project(self.pts, ax)
            b0, b1 = self._project(other.pts, ax)
            if a1 < b0 or b1 < a0:      # strictly separated on this axis
                return False
        return True


def box(minx, miny, maxx, maxy):
    return Polygon([(minx, miny), (maxx, miny), (maxx, maxy), (minx, maxy)])
