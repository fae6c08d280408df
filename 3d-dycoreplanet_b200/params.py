"""Parameter sets of the reference's named configs, and a reader for its `.prm` files.

The derived, dimensionless numbers are the ones the assembly loops actually consume
(/root/reference/include/core/boussinesq_model.tpp:564-568 `1/Re`, :760-764 `1/Pe`, :640-643 gravity
factor L/U^2, :615-621 Coriolis factor L/U only under `cuboid_geometry`), derived exactly as in
/root/reference/source/model_data/physical_constants.cc:148-164 (nu = mu/rho, kappa = k/(c_p * p_atm),
R1 = R0 + atm height) and core_model_data.cc:7-22 (Re = U L / nu, Pe = U L / kappa).

`/root/reference` is not present on the GPU box, so the named configs are tabulated here
(values cross-checked against the .prm files by tests/test_params.py when the reference tree exists).
"""
import re
from dataclasses import dataclass, asdict


@dataclass
class ModelParameters:
    # Boussinesq Model subsection (source/model_data/boussinesq_model_parameters.cc:52-239)
    space_dimension: int = 3
    initial_global_refinement: int = 2
    cuboid_geometry: bool = False
    time_step: float = 0.1
    final_time: float = 0.09
    nse_velocity_degree: int = 2
    temperature_degree: int = 1
    use_FEEC_solver: bool = False
    use_schur_complement_solver: bool = False
    NSE_solver_interval: int = 1
    # Reference quantities (source/model_data/reference_quantities.cc:37-88)
    ref_velocity: float = 1.0
    ref_length: float = 1.0
    ref_temperature: float = 2.0
    # Physical constants (source/model_data/physical_constants.cc:50-167)
    omega: float = 1.0
    density: float = 1.0
    expansion_coefficient: float = 0.2
    dynamic_viscosity: float = 1.0e-2
    specific_heat_p: float = 1.0
    thermal_conductivity: float = 1.0e-3
    pressure: float = 1.0
    gravity_constant: float = 1.0
    atm_height: float = 2.0
    R0: float = 1.0

    # ---- derived (after the model constructor's rescaling, boussinesq_model.tpp:42-63) ---------
    @property
    def kinematic_viscosity(self):
        return self.dynamic_viscosity / self.density

    @property
    def thermal_diffusivity(self):
        return self.thermal_conductivity / (self.specific_heat_p * self.pressure)

    @property
    def inv_re(self):
        return 1.0 / ((self.ref_velocity * self.ref_length) / self.kinematic_viscosity)

    @property
    def inv_pe(self):
        return 1.0 / ((self.ref_velocity * self.ref_length) / self.thermal_diffusivity)

    @property
    def R0_scaled(self):
        return self.R0 / self.ref_length

    @property
    def R1_scaled(self):
        return (self.R0 + self.atm_height) / self.ref_length

    @property
    def g_scale(self):
        return self.ref_length / (self.ref_velocity * self.ref_velocity)

    @property
    def cor_scale(self):
        return self.ref_length / self.ref_velocity

    def to_dict(self):
        return asdict(self)


# data/aqua_planet_shell_test_3d-classic.prm, ...-feec.prm, ...cube_test_3d.prm, ...test_2d.prm
NAMED = {
    "shell_3d_classic": ModelParameters(),
    "shell_3d_feec": ModelParameters(initial_global_refinement=3, use_FEEC_solver=True, nse_velocity_degree=1,
                                     expansion_coefficient=0.5, thermal_conductivity=1e-5),
    "cube_3d": ModelParameters(initial_global_refinement=4, cuboid_geometry=True, time_step=0.01, final_time=2.0,
                               use_FEEC_solver=True, use_schur_complement_solver=True, nse_velocity_degree=1,
                               ref_temperature=3.0, omega=2.0, dynamic_viscosity=1e-3),
    "annulus_2d": ModelParameters(space_dimension=2, initial_global_refinement=4, time_step=0.01, final_time=1.0,
                                  temperature_degree=2, use_schur_complement_solver=True, ref_velocity=0.1,
                                  ref_length=0.1, omega=0.5, dynamic_viscosity=1e-3,
                                  thermal_conductivity=1e-3, atm_height=2.0, R0=1.0),
}

_KEYMAP = {
    "space dimension": ("space_dimension", int), "initial global refinement": ("initial_global_refinement", int),
    "cuboid geometry": ("cuboid_geometry", "bool"), "time step": ("time_step", float),
    "final time": ("final_time", float), "nse velocity degree": ("nse_velocity_degree", int),
    "temperature degree": ("temperature_degree", int), "use FEEC solver": ("use_FEEC_solver", "bool"),
    "use schur complement solver": ("use_schur_complement_solver", "bool"),
    "NSE solver interval": ("NSE_solver_interval", int), "velocity": ("ref_velocity", float),
    "length": ("ref_length", float), "temperature": ("ref_temperature", float), "omega": ("omega", float),
    "density": ("density", float), "expansion coefficient": ("expansion_coefficient", float),
    "dynamic viscosity": ("dynamic_viscosity", float), "specific heat p": ("specific_heat_p", float),
    "thermal conductivity": ("thermal_conductivity", float), "average atm pressure": ("pressure", float),
    "gravity constant": ("gravity_constant", float), "atm height": ("atm_height", float), "R0": ("R0", float),
}


def read_prm(path):
    """Read a deal.II ParameterHandler file the way the reference's three structs do (each re-opens the same file
    with skip_undefined=true: boussinesq_model_parameters.cc:43-46, physical_constants.cc:41-44,
    reference_quantities.cc:28-31).  Only the keys the hot path consumes are kept."""
    p = ModelParameters()
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            m = re.match(r"set\s+(.+?)\s*=\s*(.+)$", line)
            if not m:
                continue
            key, val = m.group(1).strip(), m.group(2).strip()
            if key in _KEYMAP:
                name, typ = _KEYMAP[key]
                setattr(p, name, (val.lower() == "true") if typ == "bool" else typ(val))
    return p
