from . import Polygon, _Unavailable  # noqa: F401


class LinearRing(_Unavailable):
    pass
