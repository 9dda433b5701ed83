"""`AcousticDynamics` — drop-in for fv3core/pace/fv3core/stencils/dyn_core.py:220-970.

Same constructor and call signature; the body issues the same sequence of stages and halo exchanges as the
reference's `__call__` (dyn_core.py:670-970), each stage being one C-ABI call into libfv3b200.
"""
import numpy as np

from ... import constants
from ...util.quantity import Quantity
from . import acoustic_misc, d_sw as d_sw_mod, updatedzd
from .c_sw import CGridShallowWaterDynamics
from .riem_solver_c import NonhydrostaticVerticalSolverCGrid
from .updatedzc import UpdateGeopotentialHeightOnCGrid

C = constants
_C3 = (C.X_DIM, C.Y_DIM, C.Z_DIM)
_C3I = (C.X_DIM, C.Y_DIM, C.Z_INTERFACE_DIM)
_U3 = (C.X_DIM, C.Y_INTERFACE_DIM, C.Z_DIM)
_V3 = (C.X_INTERFACE_DIM, C.Y_DIM, C.Z_DIM)
_B3 = (C.X_INTERFACE_DIM, C.Y_INTERFACE_DIM, C.Z_DIM)
HUGE_R = 1.0e40


def dyncore_temporaries(quantity_factory):
    """dyn_core.py:190-218"""
    t = {}
    for name in ["ut", "vt", "pem", "pk3", "heat_source", "cappa"]:
        t[name] = quantity_factory.zeros(_C3, "unknown")
    for name in ["gz", "pkc", "zh"]:
        t[name] = quantity_factory.zeros(_C3I, "unknown")
    t["divgd"] = quantity_factory.zeros(_B3, "unknown")
    t["ws3"] = quantity_factory.zeros((C.X_DIM, C.Y_DIM), "unknown")
    for name in ["crx", "xfx"]:
        t[name] = quantity_factory.zeros(_V3, "unknown")
    for name in ["cry", "yfx"]:
        t[name] = quantity_factory.zeros(_U3, "unknown")
    return t


def get_nk_heat_dissipation(config, npz: int) -> int:
    """dyn_core.py:174-187"""
    if config.convert_ke or config.vtdm4 > 1.0e-4:
        return npz
    if config.d2_bg_k1 < 1.0e-3:
        return 0
    return 1 if config.d2_bg_k2 < 1.0e-3 else 2


class _Bound:
    """A halo updater bound to fixed quantities (the reference's WrappedHaloUpdater, wrapped_halo_exchange.py:9-73)."""

    def __init__(self, updater, qx, qy=None):
        self._u, self._qx, self._qy = updater, qx, qy

    def start(self):
        self._u.start(self._qx, self._qy)

    def wait(self):
        self._u.wait()

    def update(self):
        self._u.update(self._qx, self._qy)

    interface = update


class AcousticDynamics:
    class _HaloUpdaters:
        """dyn_core.py:227-343"""

        def __init__(self, comm, quantity_factory, state, cappa, gz, zh, divgd, heat_source, pkc):
            qf = quantity_factory
            xyz = qf.get_quantity_halo_spec(_C3)
            xyiz = qf.get_quantity_halo_spec(_U3)
            xiyz = qf.get_quantity_halo_spec(_V3)
            xyzi = qf.get_quantity_halo_spec(_C3I)
            xiyiz = qf.get_quantity_halo_spec(_B3)
            self.q_con__cappa = _Bound(comm.get_scalar_halo_updater([xyz] * 2), [state.q_con, cappa])
            self.delp__pt = _Bound(comm.get_scalar_halo_updater([xyz] * 2), [state.delp, state.pt])
            self.u__v = _Bound(comm.get_vector_halo_updater([xyiz], [xiyz]), [state.u], [state.v])
            self.w = _Bound(comm.get_scalar_halo_updater([xyz]), [state.w])
            self.gz = _Bound(comm.get_scalar_halo_updater([xyzi]), [gz])
            self.delp__pt__q_con = _Bound(comm.get_scalar_halo_updater([xyz] * 3), [state.delp, state.pt, state.q_con])
            self.zh = _Bound(comm.get_scalar_halo_updater([xyzi]), [zh])
            self.divgd = _Bound(comm.get_scalar_halo_updater([xiyiz]), [divgd])
            self.heat_source = _Bound(comm.get_scalar_halo_updater([xyz]), [heat_source])
            self.pkc = _Bound(comm.get_scalar_halo_updater([qf.get_quantity_halo_spec(_C3I, n_halo=2)]), [pkc])
            self.uc__vc = _Bound(comm.get_vector_halo_updater([xiyz], [xyiz]), [state.uc], [state.vc])
            self.interface_uc__vc = _Bound(comm.get_interface_updater(xyiz, xiyz), [state.u], [state.v])

    def __init__(self, comm, stencil_factory, quantity_factory, grid_data, damping_coefficients, grid_type, nested,
                 stretched_grid, config, phis: Quantity, wsd: Quantity, state, checkpointer=None):
        if config.d_ext != 0:
            raise NotImplementedError("d_ext != 0 is not implemented")
        if config.beta != 0:
            raise NotImplementedError("beta != 0 is not implemented")
        if config.use_logp:
            raise NotImplementedError("use_logp=True is not implemented")
        if config.hydrostatic:
            raise NotImplementedError("hydrostatic dynamics are not implemented")
        self._rt = rt = stencil_factory.runtime
        self.config = config
        self.grid_data = grid_data
        self.checkpointer = checkpointer
        self.call_checkpointer = checkpointer is not None
        self._da_min = damping_coefficients.da_min
        self._ptop = grid_data.ptop
        self._wsd = wsd
        nz = rt.comm.geometry.nz
        self._nk_heat_dissipation = get_nk_heat_dissipation(config.d_grid_shallow_water, nz)
        t = dyncore_temporaries(quantity_factory)
        self._heat_source, self._divgd, self._gz, self._pkc, self._zh = t["heat_source"], t["divgd"], t["gz"], t["pkc"], t["zh"]
        self.cappa, self._ut, self._vt, self._pem, self._pk3 = t["cappa"], t["ut"], t["vt"], t["pem"], t["pk3"]
        self._crx, self._cry, self._xfx, self._yfx, self._ws3 = t["crx"], t["cry"], t["xfx"], t["yfx"], t["ws3"]
        self._pk3.data[:] = HUGE_R
        self._zs = quantity_factory.zeros((C.X_DIM, C.Y_DIM), "m")
        self._zs.data[:] = phis.data / C.GRAV
        cols = d_sw_mod.ColumnNamelist(rt, config.d_grid_shallow_water, damping_coefficients)
        self._column_namelist = cols
        self.update_height_on_d_grid = updatedzd.UpdateHeightOnDGrid(
            stencil_factory, quantity_factory, damping_coefficients, grid_data, grid_type, config.hord_tm, cols)
        self.vertical_solver = acoustic_misc.NonhydrostaticVerticalSolver(stencil_factory, quantity_factory, config.riemann)
        self.vertical_solver_cgrid = NonhydrostaticVerticalSolverCGrid(stencil_factory, quantity_factory, config.p_fac)
        self.dgrid_shallow_water_lagrangian_dynamics = d_sw_mod.DGridShallowWaterLagrangianDynamics(
            stencil_factory, quantity_factory, grid_data, damping_coefficients, cols, nested, stretched_grid,
            config.d_grid_shallow_water)
        self.cgrid_shallow_water_lagrangian_dynamics = CGridShallowWaterDynamics(
            stencil_factory, quantity_factory, grid_data, nested, config.grid_type, config.nord)
        self.update_geopotential_height_on_c_grid = UpdateGeopotentialHeightOnCGrid(
            stencil_factory, quantity_factory, grid_data.area, grid_data.dp_ref)
        self.nonhydrostatic_pressure_gradient = acoustic_misc.NonHydrostaticPressureGradient(
            stencil_factory, quantity_factory, grid_data, config.grid_type)
        self._do_del2cubed = self._nk_heat_dissipation != 0 and config.d_con > 1.0e-5
        if self._do_del2cubed:
            self._hyperdiffusion = acoustic_misc.HyperdiffusionDamping(
                stencil_factory, quantity_factory, damping_coefficients, grid_data.rarea, min(3, config.nord + 1))
        if config.rf_fast:
            self._rayleigh_damping = acoustic_misc.RayleighDamping(stencil_factory, config.rf_cutoff, config.tau, config.hydrostatic)
        self._pk3_halo = acoustic_misc.PK3Halo(stencil_factory, quantity_factory)
        self._halo_updaters = AcousticDynamics._HaloUpdaters(
            comm, quantity_factory, state, self.cappa, self._gz, self._zh, self._divgd, self._heat_source, self._pkc)

    def _zero_data(self, state, first_timestep: bool):
        """zero_data (dyn_core.py:48-80)"""
        import torch

        fields = [state.mfxd, state.mfyd, state.cxd, state.cyd]
        if first_timestep:  # domain_full, as the reference
            fields += [self._heat_source, state.diss_estd]
        torch._foreach_zero_([q.data for q in fields])   # one multi-tensor launch instead of one fill per field

    def __call__(self, state, timestep: float, n_map=1):
        rt, cfg, hu = self._rt, self.config, self._halo_updaters
        end_step = n_map == cfg.k_split
        akap = C.KAPPA
        dt = timestep / cfg.n_split
        dt2 = 0.5 * dt
        n_split = cfg.n_split
        csw = self.cgrid_shallow_water_lagrangian_dynamics
        hu.q_con__cappa.start()
        hu.delp__pt.start()
        hu.u__v.start()
        hu.q_con__cappa.wait()
        self._zero_data(state, n_map == 1)
        for it in range(n_split):
            remap_step = cfg.breed_vortex_inline or (it == n_split - 1)
            hu.w.start()
            if it == 0:
                rt.call("fv3_gz_from_delz", self._zs.ptr, state.delz.ptr, self._gz.ptr)
                hu.gz.start()
                hu.delp__pt.wait()
            if it == n_split - 1 and end_step and cfg.use_old_omega:
                rt.call("fv3_pem_from_delp", state.delp.ptr, self._pem.ptr, float(self._ptop))
            hu.u__v.wait()
            hu.w.wait()
            if self.call_checkpointer:
                self._checkpoint_csw(state, "In")
            csw(state.delp, state.pt, state.u, state.v, state.w, state.uc, state.vc, state.ua, state.va, self._ut, self._vt,
                self._divgd, state.omga, dt2)
            if self.call_checkpointer:
                self._checkpoint_csw(state, "Out")
            if cfg.nord > 0:
                hu.divgd.start()
            if it == 0:
                hu.gz.wait()
                self._zh.data.copy_(self._gz.data)
            else:
                self._gz.data.copy_(self._zh.data)
            self.update_geopotential_height_on_c_grid(self._zs, self._ut, self._vt, self._gz, self._ws3, dt2)
            self.vertical_solver_cgrid(dt2, self.cappa, self._ptop, state.phis, self._ws3, csw.ptc, state.q_con, csw.delpc,
                                       self._gz, self._pkc, state.omga)
            rt.call("fv3_p_grad_c", self.grid_data.rdxc.ptr, self.grid_data.rdyc.ptr, state.uc.ptr, state.vc.ptr,
                    csw.delpc.ptr, self._pkc.ptr, self._gz.ptr, float(dt2))
            hu.uc__vc.start()
            if cfg.nord > 0:
                hu.divgd.wait()
            hu.uc__vc.wait()
            if self.call_checkpointer:
                self._checkpoint_dsw(state, "D_SW-In")
            self.dgrid_shallow_water_lagrangian_dynamics(
                self._vt, state.delp, state.pt, state.u, state.v, state.w, state.uc, state.vc, state.ua, state.va,
                self._divgd, state.mfxd, state.mfyd, state.cxd, state.cyd, self._crx, self._cry, self._xfx, self._yfx,
                state.q_con, self._zh, self._heat_source, state.diss_estd, dt)
            if self.call_checkpointer:
                self._checkpoint_dsw(state, "D_SW-Out")
            hu.delp__pt__q_con.update()
            self.update_height_on_d_grid(self._zs, self._zh, self._crx, self._cry, self._xfx, self._yfx, self._wsd, dt)
            self.vertical_solver(remap_step, dt, self.cappa, self._ptop, self._zs, self._wsd, state.delz, state.q_con,
                                 state.delp, state.pt, self._zh, state.pe, self._pkc, self._pk3, state.pk, state.peln,
                                 state.w)
            hu.zh.start()
            hu.pkc.start()
            if remap_step:
                rt.call("fv3_edge_pe", state.pe.ptr, state.delp.ptr, float(self._ptop))
            self._pk3_halo(self._pk3, state.delp, self._ptop, akap)
            hu.zh.wait()
            rt.call("fv3_compute_geopotential", self._zh.ptr, self._gz.ptr)
            hu.pkc.wait()
            self.nonhydrostatic_pressure_gradient(state.u, state.v, self._pkc, self._gz, self._pk3, state.delp, dt,
                                                  self._ptop, akap)
            if cfg.rf_fast:
                self._rayleigh_damping(state.u, state.v, state.w, None, None, dt, self._ptop)
            if it != n_split - 1:
                hu.u__v.start()
            elif cfg.grid_type < 4:
                hu.interface_uc__vc.interface()
        if self._do_del2cubed:
            hu.heat_source.update()
            cd = C.CNST_0P20 * self._da_min
            self._hyperdiffusion(self._heat_source, cd)
            delt_time_factor = abs(dt * cfg.delt_max)
            rt.call("fv3_apply_diffusive_heating", state.delp.ptr, state.delz.ptr, self.cappa.ptr,
                    self._heat_source.ptr, state.pt.ptr, float(delt_time_factor), int(self._nk_heat_dissipation))

    # -- checkpoint hooks (dyn_core.py:608-668): same names and variables as the reference ------------------
    def _checkpoint_csw(self, state, tag):
        self.checkpointer(f"C_SW-{tag}", delpd=state.delp, ptd=state.pt, ud=state.u, vd=state.v, wd=state.w, ucd=state.uc,
                          vcd=state.vc, uad=state.ua, vad=state.va, utd=self._ut, vtd=self._vt, divgdd=self._divgd)

    def _checkpoint_dsw(self, state, name):
        self.checkpointer(name, ucd=state.uc, vcd=state.vc, wd=state.w, delpcd=self._vt, delpd=state.delp, ud=state.u,
                          vd=state.v, ptd=state.pt, uad=state.ua, vad=state.va, divgdd=self._divgd, xfxd=self._xfx,
                          yfxd=self._yfx, mfxd=state.mfxd, mfyd=state.mfyd)
