"""
Fixtures mirroring the reference's tests/pf/test_forward_Grav_Linear.py:423-456
and tests/pf/test_forward_Mag_Linear.py:915-975: 4x4x4 mesh (dh=5, "CCN"),
4x4 receivers at z=5, sphere model.
"""
import numpy as np

from oracle.oracle import OracleSimulation
from simpeg_b200 import maps
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.potential_fields import gravity, magnetics


def small_mesh():
    return TensorMesh([[(5.0, 4)], [(5.0, 4)], [(5.0, 4)]], "CCN")


def receiver_locations():
    x = np.linspace(-20.0, 20.0, 4)
    x, y = np.meshgrid(x, x)
    z = 5.0 * np.ones_like(x)
    return np.vstack((x.ravel(), y.ravel(), z.ravel())).T


def sphere_indices(mesh, center=(0.0, 0.0, -20.0), radius=10.0):
    return np.linalg.norm(mesh.cell_centers - np.asarray(center), axis=1) < radius


def gravity_survey(components="gz"):
    rx = gravity.receivers.Point(receiver_locations(), components=components)
    return gravity.survey.Survey(gravity.sources.SourceField(receiver_list=[rx]))


def densities(mesh):
    rho = 1e-10 * np.ones(mesh.n_cells)
    rho[sphere_indices(mesh)] = 0.2
    return rho


def magnetic_survey(components="tmi", two_objects=False):
    loc = receiver_locations()
    if two_objects:
        rxs = [magnetics.receivers.Point(loc[:7], components=components),
               magnetics.receivers.Point(loc[7:], components=["bz", "tmi"])]
    else:
        rxs = [magnetics.receivers.Point(loc, components=components)]
    src = magnetics.sources.UniformBackgroundField(
        receiver_list=rxs, amplitude=55_000, inclination=12, declination=-35)
    return magnetics.survey.Survey(src)


def susceptibilities(mesh, model_type):
    chi = 1e-10 * np.ones(mesh.n_cells)
    chi[sphere_indices(mesh)] = 0.2
    if model_type == "vector":
        rng = np.random.default_rng(7)
        chi = np.concatenate([chi, 0.5 * chi, -0.3 * chi]) + 1e-3 * rng.uniform(size=3 * mesh.n_cells)
    return chi


def make_mapping(kind, n):
    return maps.IdentityMap(nP=n) if kind == "identity_map" else maps.ExpMap(nP=n)


def oracle_for(sim):
    """OracleSimulation (C restatement of the reference loops) for a simulation."""
    nodes, cell_nodes = sim._get_active_nodes()
    rlist = [(loc, comps) for comps, loc in sim._get_components_and_receivers()]
    b0 = None
    model_type = "scalar"
    physics = sim._physics
    if physics == "magnetics":
        b0 = sim.survey.source_field.b0
        model_type = sim.model_type
    return OracleSimulation(physics, nodes, cell_nodes, rlist, b0=b0, model_type=model_type)
