def make_surface(arr):
    from . import Surface
    return Surface(arr.shape[:2])
