"""waterlily_b200 — B200-native (sm_100a CUDA) implementation of WaterLily.jl's ``mom_step!`` path
behind a C ABI, with a host-side mirror of the reference's user API.

    from waterlily_b200 import Simulation, sim_step, sim_time, AutoBody
    sim = Simulation((128,128,128), (0,0,0), 128, U=1, nu=..., perdir=(1,2,3), ulam=tgv)
    sim_step(sim, 1.0, remeasure=False)
    u = sim.flow.u.numpy()
"""
from . import _cabi
from ._cabi import Engine, WLError
from .body import AbstractBody, AutoBody, NoBody
from .sim import (BCTuple, DeviceField, Flow, MultiLevelPoisson, _SimulationBase, curl_field, ke_field, lambda2_field, measure,
                  measure_flow, mom_step, omega_mag_field, omega_theta_field, pressure_force, pressure_moment, sim_step, sim_time,
                  time, total_force, viscous_force)


class Simulation(_SimulationBase):
    """``Simulation(dims,u_BC,L;...)`` on one B200 (reference src/WaterLily.jl:59-79, ``mem=CuArray``)."""

    @classmethod
    def _lib(cls):
        from .lib import load
        return load()


__all__ = ["Simulation", "sim_step", "sim_time", "time", "measure", "mom_step", "Flow", "MultiLevelPoisson",
           "AbstractBody", "AutoBody", "NoBody", "BCTuple", "Engine", "WLError", "DeviceField",
           "pressure_force", "viscous_force", "total_force", "pressure_moment",
           "ke_field", "curl_field", "omega_mag_field", "lambda2_field", "omega_theta_field"]
