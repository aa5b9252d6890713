"""Minimal stand-in for pygame 2.1.2 (test infrastructure; see ../README.md).

Only `Surface.get_rect(center=...)`, `Rect` corners and `math.Vector2` carry
arithmetic on the reference's hot path (merging_env.py:97-98, 232-239).
"""
from . import math, locals, surfarray, display, font, time, draw, event, image, transform, key  # noqa: F401,E501


def _c_int(v):
    # pygame 2.1.2 pg_IntFromObj: (int)PyFloat_AsDouble(obj) -> truncation toward zero.
    return int(v)


class Rect:
    def __init__(self, x, y, w, h):
        self.x, self.y, self.w, self.h = _c_int(x), _c_int(y), _c_int(w), _c_int(h)

    # the `center` setter moves the rect so that x + (w >> 1) == int(cx)
    def _set_center(self, c):
        cx, cy = _c_int(c[0]), _c_int(c[1])
        self.x += cx - (self.x + (self.w >> 1))
        self.y += cy - (self.y + (self.h >> 1))

    center = property(lambda s: (s.x + (s.w >> 1), s.y + (s.h >> 1)), _set_center)
    topleft = property(lambda s: (s.x, s.y))
    topright = property(lambda s: (s.x + s.w, s.y))
    bottomleft = property(lambda s: (s.x, s.y + s.h))
    bottomright = property(lambda s: (s.x + s.w, s.y + s.h))
    width = property(lambda s: s.w)
    height = property(lambda s: s.h)


class Surface:
    def __init__(self, size, *a, **k):
        self._w, self._h = int(size[0]), int(size[1])

    def get_rect(self, **kwargs):
        r = Rect(0, 0, self._w, self._h)
        for name, val in kwargs.items():
            setattr(r, name, val)
        return r

    def get_size(self):
        return (self._w, self._h)

    def fill(self, *a, **k):
        pass

    def blit(self, *a, **k):
        pass

    def convert(self, *a, **k):
        return self

    def set_colorkey(self, *a, **k):
        pass


def init():
    return (0, 0)


def quit():
    pass
