class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("shapely is not installed in this image; only the training hot path is runnable")


class MultiLineString(_Unavailable):
    pass


class Polygon(_Unavailable):
    pass


class JOIN_STYLE:
    round = 1
    mitre = 2
    bevel = 3
