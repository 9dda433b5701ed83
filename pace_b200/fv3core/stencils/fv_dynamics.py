"""`DynamicalCore` — drop-in for fv3core/pace/fv3core/stencils/fv_dynamics.py:92-624.

Same constructor, `step_dynamics(state, timer)` / `__call__`, attributes `acoustic_dynamics`,
`tracer_advection`, `_lagrangian_to_eulerian_obj`.  One instance drives every subdomain resident on this GPU.
"""
from datetime import timedelta

import torch

from ... import constants as c
from ...util.timer import NullTimer
from ..dycore_state import TRACER_VARIABLES, DycoreState
from . import acoustic_misc, fvtp2d
from .dyn_core import AcousticDynamics
from .neg_adj3 import AdjustNegativeTracerMixingRatio
from .remapping import LagrangianToEulerian
from .tracer_2d_1l import TracerAdvection

NQ = 8  # fv_dynamics.py:36
_C3 = (c.X_DIM, c.Y_DIM, c.Z_DIM)
_U3 = (c.X_DIM, c.Y_INTERFACE_DIM, c.Z_DIM)
_V3 = (c.X_INTERFACE_DIM, c.Y_DIM, c.Z_DIM)


class DynamicalCore:
    def __init__(self, comm, grid_data, stencil_factory, quantity_factory, damping_coefficients, config, phis,
                 state: DycoreState, timestep: timedelta, checkpointer=None):
        if not config.moist_phys:
            raise NotImplementedError("fvsetup is only implemented for moist_phys=true")
        if config.nwat != 6:
            raise NotImplementedError("Only nwat=6 has been implemented and tested")
        if config.hydrostatic:
            raise NotImplementedError("Hydrostatic is not implemented")
        if config.inline_q or NQ == 0:
            raise NotImplementedError("tracer_2d not implemented, turn on z_tracer")
        if not config.z_tracer:
            raise NotImplementedError("z_tracer=False is not implemented")
        if (not config.rf_fast) and config.tau != 0:
            raise NotImplementedError("Rayleigh_Super, called when rf_fast=False and tau !=0")
        if config.consv_te > 0:
            raise NotImplementedError("compute total energy is not implemented")
        if config.adiabatic and config.kord_tm > 0:
            raise NotImplementedError("unimplemented namelist options adiabatic with positive kord_tm")
        if config.c2l_ord != 4:
            raise NotImplementedError("only c2l_ord=4 is implemented")
        self._rt = rt = stencil_factory.runtime
        self.checkpointer = checkpointer
        self.call_checkpointer = checkpointer is not None
        self.comm = comm
        self.comm_rank = comm.rank
        self.grid_data = grid_data
        self.config = config
        self._da_min = damping_coefficients.da_min
        qf = quantity_factory
        tracer_transport = fvtp2d.FiniteVolumeTransport(stencil_factory, qf, grid_data, damping_coefficients,
                                                        config.grid_type, config.hord_tr)
        self.tracers = {name: getattr(state, name) for name in TRACER_VARIABLES[:NQ]}
        self._wsd = qf.zeros((c.X_DIM, c.Y_DIM), "unknown")
        self._dp_initial = qf.zeros(_C3, "unknown")
        self._cvm = qf.zeros(_C3, "unknown")
        self.tracer_advection = TracerAdvection(stencil_factory, qf, tracer_transport, grid_data, comm, self.tracers)
        self._phis = phis
        self._ptop = grid_data.ptop
        self.acoustic_dynamics = AcousticDynamics(
            comm, stencil_factory, qf, grid_data, damping_coefficients, config.grid_type, False, False,
            config.acoustic_dynamics, phis, self._wsd, state, checkpointer)
        self._hyperdiffusion = acoustic_misc.HyperdiffusionDamping(stencil_factory, qf, damping_coefficients,
                                                                   grid_data.rarea, config.nf_omega)
        self._cappa = self.acoustic_dynamics.cappa
        self._adjust_tracer_mixing_ratio = AdjustNegativeTracerMixingRatio(
            stencil_factory, quantity_factory=qf, check_negative=getattr(config, "check_negative", False),
            hydrostatic=config.hydrostatic)
        self._lagrangian_to_eulerian_obj = LagrangianToEulerian(
            stencil_factory, qf, config.remapping, grid_data.area_64, NQ, None, self.tracers, checkpointer)
        self._omega_halo_updater = comm.get_scalar_halo_updater([qf.get_quantity_halo_spec(_C3)])
        self._c2l_updater = comm.get_vector_halo_updater([qf.get_quantity_halo_spec(_U3)], [qf.get_quantity_halo_spec(_V3)])
        names6 = ["qvapor", "qliquid", "qrain", "qsnow", "qice", "qgraupel"]
        self._t6 = torch.tensor([getattr(state, n).ptr for n in names6], dtype=torch.int64).to(rt.device)
        self._bound_state = state
        self._bound_ptrs = {n: getattr(state, n).ptr for n in list(TRACER_VARIABLES[:NQ]) + ["u", "v", "w", "delp", "pt", "delz", "q_con"]}
        self._n_split, self._k_split = config.n_split, config.k_split
        self._timestep = timestep.total_seconds()

    def _checkpoint_fvdynamics(self, state, tag):
        if self.call_checkpointer:
            self.checkpointer(f"FVDynamics-{tag}", u=state.u, v=state.v, w=state.w, delz=state.delz, va=state.va,
                              uc=state.uc, vc=state.vc, qvapor=state.qvapor)

    def _checkpoint_remapping(self, state, tag):
        """Remapping-In / Remapping-Out (fv_dynamics.py:340-395); te_2d is left out (consv_te = 0 never fills it), pe and
        peln are handed over in this repo's (i, j, k) order, not the Fortran (i, k, j) the reference transposes to."""
        if not self.call_checkpointer:
            return
        common = dict(pt=state.pt, delp=state.delp, delz=state.delz, peln=state.peln, u=state.u, v=state.v, w=state.w,
                      cappa=self._cappa, pk=state.pk, pe=state.pe, dp1=self._dp_initial)
        if tag == "In":
            self.checkpointer("Remapping-In", ua=state.ua, va=state.va, phis=state.phis, ps=state.ps, wsd=self._wsd,
                              omga=state.omga, **common)
        else:
            self.checkpointer("Remapping-Out", pkz=state.pkz, **common)

    def _check_bound_state(self, state):
        """The tracer pointer table, the tracer dictionary and every halo updater are bound to the DycoreState given at
        construction (the reference binds its halo updaters the same way, dyn_core.py:273-343).  A different state
        object must not silently advance with the construction-time tracers."""
        if state is self._bound_state:
            return
        for n in self._bound_ptrs:
            if getattr(state, n).ptr != self._bound_ptrs[n]:
                raise ValueError(f"step_dynamics: state.{n} is not the buffer this DynamicalCore was constructed with; "
                                 "the halo updaters and tracer tables are bound at construction")

    def step_dynamics(self, state: DycoreState, timer=NullTimer()):
        self._check_bound_state(state)
        self._checkpoint_fvdynamics(state, "In")
        self._compute(state, timer)
        self._checkpoint_fvdynamics(state, "Out")

    __call__ = step_dynamics

    def capture_step(self, state: DycoreState, warmup: int = 1):
        """Capture one `step_dynamics(state)` — every kernel launch and halo exchange of the timestep — into a CUDA
        graph and return `replay()`, which advances `state` in place by one timestep per call.

        B200-first replacement of the reference's DaCe orchestration (dsl/pace/dsl/dace/orchestration.py): the step is
        ~1600 short launches, so on a GPU that owns only a few subdomains the host enqueue time would otherwise bound
        the step.  The stage sequence is static (n_split, k_split fixed at construction), which is what makes it
        capturable.  The checkpointer must be off."""
        if self.call_checkpointer:
            raise RuntimeError("capture_step cannot be used with a checkpointer")
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step_dynamics(state)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step_dynamics(state)
        self._graph = graph
        return graph.replay

    def compute_preamble(self, state, is_root_rank: bool = True):
        self._rt.call("fv3_fv_setup", self._t6.data_ptr(), state.q_con.ptr, self._cvm.ptr, state.pkz.ptr, state.pt.ptr,
                      self._cappa.ptr, state.delp.ptr, state.delz.ptr, self._dp_initial.ptr)

    def _compute(self, state, timer):
        rt = self._rt
        self.compute_preamble(state, self.comm_rank == 0)
        for k_split in range(self._k_split):
            n_map = k_split + 1
            last_step = k_split == self._k_split - 1
            self._dp_initial.data.copy_(state.delp.data)
            with timer.clock("DynCore"):
                self.acoustic_dynamics(state, timestep=self._timestep / self._k_split, n_map=n_map)
            with timer.clock("TracerAdvection"):
                if self.call_checkpointer:
                    self.checkpointer("Tracer2D1L-In", dp1=self._dp_initial, mfxd=state.mfxd, mfyd=state.mfyd,
                                      cxd=state.cxd, cyd=state.cyd)
                self.tracer_advection(self.tracers, self._dp_initial, state.mfxd, state.mfyd, state.cxd, state.cyd)
                if self.call_checkpointer:
                    self.checkpointer("Tracer2D1L-Out", dp1=self._dp_initial, mfxd=state.mfxd, mfyd=state.mfyd,
                                      cxd=state.cxd, cyd=state.cyd)
            if rt.comm.geometry.nz > 4:
                with timer.clock("Remapping"):
                    self._checkpoint_remapping(state, "In")
                    self._lagrangian_to_eulerian_obj(
                        self.tracers, state.pt, state.delp, state.delz, state.peln, state.u, state.v, state.w,
                        self._cappa, state.q_con, state.qcld, state.pkz, state.pk, state.pe, state.phis, state.ps,
                        self._wsd, None, None, self._dp_initial, self._ptop, c.KAPPA, c.ZVIR, last_step,
                        self.config.consv_te, self._timestep / self._k_split)
                    self._checkpoint_remapping(state, "Out")
                if last_step:
                    rt.call("fv3_omega_from_w", state.delp.ptr, state.delz.ptr, state.w.ptr, state.omga.ptr)
                    if self.config.nf_omega > 0:
                        self._omega_halo_updater.update([state.omga])
                        self._hyperdiffusion(state.omga, 0.18 * self._da_min)
        self._adjust_tracer_mixing_ratio(state.qvapor, state.qliquid, state.qrain, state.qsnow, state.qice, state.qgraupel,
                                         state.qcld, state.pt, state.delp)
        self._c2l_updater.update([state.u], [state.v])
        rt.call("fv3_c2l_ord4", state.u.ptr, state.v.ptr, state.ua.ptr, state.va.ptr)
