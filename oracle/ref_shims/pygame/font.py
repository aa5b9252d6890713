class Font:
    def __init__(self, *a, **k):
        pass

    def render(self, *a, **k):
        from . import Surface
        return Surface((1, 1))


def SysFont(*a, **k):
    return Font()


def init():
    pass
