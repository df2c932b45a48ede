"""gravity/receivers.py:5-85 mirror: point receivers with component lists."""
from .._survey import BasePoint


class Point(BasePoint):
    _allowed = ("gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "guv")
    _default = "gz"
