"""magnetics/sources.py:8-109 mirror: uniform inducing field."""
from ...utils import dip_azimuth2cartesian


class UniformBackgroundField:
    def __init__(self, receiver_list=None, amplitude=50000.0, inclination=90.0,
                 declination=0.0, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self.receiver_list = [] if receiver_list is None else (
            [receiver_list] if not isinstance(receiver_list, (list, tuple)) else list(receiver_list))
        self.amplitude = float(amplitude)
        self.inclination = float(inclination)
        self.declination = float(declination)

    @property
    def b0(self):
        # magnetics/sources.py:104-109
        return self.amplitude * dip_azimuth2cartesian(self.inclination, self.declination).squeeze()


SourceField = UniformBackgroundField
