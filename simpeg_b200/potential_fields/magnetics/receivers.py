"""magnetics/receivers.py:5-75 mirror: point receivers with component lists."""
from .._survey import BasePoint


class Point(BasePoint):
    _allowed = ("bxx", "bxy", "bxz", "byy", "byz", "bzz", "bx", "by", "bz", "tmi",
                "tmi_x", "tmi_y", "tmi_z")
    _default = "tmi"
