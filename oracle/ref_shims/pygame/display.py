def set_mode(size, *a, **k):
    from . import Surface
    return Surface(size)


def set_caption(*a, **k):
    pass


def flip():
    pass


def update(*a, **k):
    pass
