import math as _m


class Vector2:
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=None):
        if y is None:
            if isinstance(x, Vector2):
                x, y = x.x, x.y
            elif hasattr(x, "__len__"):
                x, y = x[0], x[1]
            else:
                y = x
        self.x, self.y = float(x), float(y)

    def __sub__(self, o):
        return Vector2(self.x - o.x, self.y - o.y)

    def __add__(self, o):
        return Vector2(self.x + o.x, self.y + o.y)

    def __mul__(self, s):
        return Vector2(self.x * s, self.y * s)

    __rmul__ = __mul__

    def rotate(self, angle):
        # pygame special-cases multiples of 90 degrees; angle == 0 returns an exact copy.
        a = angle % 360.0
        if a == 0.0:
            return Vector2(self.x, self.y)
        r = _m.radians(a)
        c, s = _m.cos(r), _m.sin(r)
        return Vector2(c * self.x - s * self.y, s * self.x + c * self.y)

    def __iter__(self):
        yield self.x
        yield self.y
