"""Helpers for the equivalent-source layer tests (golden fixtures under tests/golden/layer)."""
from pathlib import Path

import numpy as np

from simpeg_b200 import maps
from simpeg_b200.meshes import CellBoundsMesh2D
from simpeg_b200.potential_fields import gravity, magnetics

ROOT = Path(__file__).resolve().parent.parent
LAYER_GOLDEN = sorted(p.stem for p in (ROOT / "tests" / "golden" / "layer").glob("*.npz"))


def load_layer_golden(name):
    from oracle.make_golden_layer import load_case

    return load_case(ROOT / "tests" / "golden" / "layer" / f"{name}.npz")


class _FixedField(magnetics.sources.UniformBackgroundField):
    """Inducing field with exactly the vector of the fixture (no angle round trip)."""

    @property
    def b0(self):
        return self._b0


def layer_simulation(case, store="ram", dtype=np.float64, mesh=None):
    """``SimulationEquivalentSourceLayer(engine="cuda")`` for a golden case."""
    mesh = CellBoundsMesh2D(case["bounds"]) if mesh is None else mesh
    nC = case["bounds"].shape[0]
    if case["physics"] == "gravity":
        rxs = [gravity.receivers.Point(loc, components=comps) for loc, comps in case["rlist"]]
        survey = gravity.survey.Survey(gravity.sources.SourceField(receiver_list=rxs))
        return gravity.SimulationEquivalentSourceLayer(
            mesh, case["top"], case["bottom"], survey=survey, rhoMap=maps.IdentityMap(nP=nC),
            store_sensitivities=store, sensitivity_dtype=dtype)
    rxs = [magnetics.receivers.Point(loc, components=comps) for loc, comps in case["rlist"]]
    b0 = np.asarray(case["b0"], dtype=np.float64)
    amp = float(np.linalg.norm(b0))
    inc = float(np.rad2deg(np.arcsin(-b0[2] / amp)))
    dec = float(np.rad2deg(np.arctan2(b0[0], b0[1])))
    src = _FixedField(receiver_list=rxs, amplitude=amp, inclination=inc, declination=dec)
    src._b0 = b0
    survey = magnetics.survey.Survey(src)
    nP = nC * (3 if case["model_type"] == "vector" else 1)
    return magnetics.SimulationEquivalentSourceLayer(
        mesh, case["top"], case["bottom"], survey=survey, chiMap=maps.IdentityMap(nP=nP),
        model_type=case["model_type"], store_sensitivities=store, sensitivity_dtype=dtype)
