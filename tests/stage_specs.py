"""Registry of hot-path stages: golden file, oracle call, native call, outputs, tolerance.

Each spec is used three ways by tests/test_stages.py: oracle vs reference golden, native kernels (host
simulation on CPU boxes / CUDA on the GPU box) vs golden, and native vs oracle.
"""
import dataclasses
from typing import Callable, Dict, Sequence

import numpy as np

from oracle.indexing import Idx
from tests import helpers as H

NX = 12
NZ = 79


@dataclasses.dataclass
class StageSpec:
    name: str
    golden: str                       # stage file name, e.g. "C_SW#0"
    outputs: Sequence[str]
    oracle: Callable                  # (ix, grid: dict, a: dict of numpy arrays) -> None (in place; may add keys)
    native: Callable                  # (sf, qf, rt, q: dict of Quantity, d: golden dict) -> None (may add keys)
    tol: float = 1e-14                # reference metric, native vs golden
    near_zero: float = 1e-18
    oracle_tol: float = 1e-14
    interface_fields: Sequence[str] = ()
    regions: Dict[str, tuple] = dataclasses.field(default_factory=dict)  # output -> (i slice, j slice) compared
    check_untouched: bool = True
    expected: Callable = None         # optional hook: golden dict -> golden dict with adjusted "out." entries
    tols: Dict[str, float] = dataclasses.field(default_factory=dict)  # per-output override of tol
    case: str = "c12"                 # golden case directory (c12 = first step, c12s2 = second step of the same run)


SPECS: Dict[str, StageSpec] = {}


def register(spec: StageSpec):
    SPECS[spec.name] = spec
    return spec


def f(d, key):
    return float(d["in." + key])


# ---------------------------------------------------------------------------------------------
def _o_riem_c(ix, g, a):
    from oracle import riem_solver as O

    O.riem_solver_c(float(a["dt2"]), a["cappa"], float(a["ptop"]), a["hs"], a["ws"], a["ptc"], a["q_con"], a["delpc"],
                    a["gz"], a["pef"], a["w3"], 0.05, ix.nx, ix.ny, ix.nz)


def _n_riem_c(sf, qf, rt, q, d):
    from pace_b200.fv3core.stencils.riem_solver_c import NonhydrostaticVerticalSolverCGrid

    NonhydrostaticVerticalSolverCGrid(sf, qf, 0.05)(f(d, "dt2"), q["cappa"], f(d, "ptop"), q["hs"], q["ws"], q["ptc"],
                                                    q["q_con"], q["delpc"], q["gz"], q["pef"], q["w3"])


register(StageSpec("riem_solver_c", "Riem_Solver_C#0", ("gz", "pef"), _o_riem_c, _n_riem_c, tol=1e-12))


def _o_c_sw(ix, g, a):
    from oracle import c_sw as O

    a["delpc"], a["ptc"] = O.c_sw(ix, g, a["delp"], a["pt"], a["u"], a["v"], a["w"], a["uc"], a["vc"], a["ua"], a["va"],
                                  a["ut"], a["vt"], a["divgd"], a["omga"], float(a["dt2"]))


def _n_c_sw(sf, qf, rt, q, d):
    from pace_b200.fv3core.stencils.c_sw import CGridShallowWaterDynamics

    csw = CGridShallowWaterDynamics(sf, qf, rt.grid_data, False, 0, 3)
    q["delpc"], q["ptc"] = csw(q["delp"], q["pt"], q["u"], q["v"], q["w"], q["uc"], q["vc"], q["ua"], q["va"], q["ut"],
                               q["vt"], q["divgd"], q["omga"], f(d, "dt2"))


register(StageSpec("c_sw", "C_SW#0", ("delp", "pt", "w", "uc", "vc", "ua", "va", "ut", "vt", "divgd", "omga", "delpc", "ptc"),
                   _o_c_sw, _n_c_sw))


def _o_dzc(ix, g, a):
    from oracle import updatedz as O

    O.update_dz_c(ix, g["dp_ref"], a["zs"], g["area"], a["ut"], a["vt"], a["gz"], a["ws"], float(a["dt"]))


def _n_dzc(sf, qf, rt, q, d):
    from pace_b200.fv3core.stencils.updatedzc import UpdateGeopotentialHeightOnCGrid

    UpdateGeopotentialHeightOnCGrid(sf, qf, rt.grid_data.area, rt.grid_data.dp_ref)(q["zs"], q["ut"], q["vt"], q["gz"],
                                                                                      q["ws"], f(d, "dt"))


register(StageSpec("update_dz_c", "UpdateDzC#0", ("gz", "ws"), _o_dzc, _n_dzc))


def _o_pgc(ix, g, a):
    from oracle import dyn_core as O

    O.p_grad_c(ix, a["rdxc"], a["rdyc"], a["uc"], a["vc"], a["delpc"], a["pkc"], a["gz"], float(a["dt2"]))


def _n_pgc(sf, qf, rt, q, d):
    rt.call("fv3_p_grad_c", q["rdxc"].ptr, q["rdyc"].ptr, q["uc"].ptr, q["vc"].ptr, q["delpc"].ptr, q["pkc"].ptr,
            q["gz"].ptr, f(d, "dt2"))


register(StageSpec("p_grad_c", "PGradC#0", ("uc", "vc"), _o_pgc, _n_pgc))


def _o_gzdelz(ix, g, a):
    from oracle import dyn_core as O

    O.gz_from_surface_height_and_thicknesses(ix, a["zs"], a["delz"], a["gz"])


register(StageSpec("gz_from_delz", "GzFromDelz#0", ("gz",), _o_gzdelz,
                   lambda sf, qf, rt, q, d: rt.call("fv3_gz_from_delz", q["zs"].ptr, q["delz"].ptr, q["gz"].ptr)))


def _o_pem(ix, g, a):
    from oracle import dyn_core as O

    O.interface_pressure_from_toa_pressure_and_thickness(ix, a["delp"], a["pem"], float(a["ptop"]))


register(StageSpec("pem_from_delp", "PemFromDelp#0", ("pem",), _o_pem,
                   lambda sf, qf, rt, q, d: rt.call("fv3_pem_from_delp", q["delp"].ptr, q["pem"].ptr, f(d, "ptop"))))


def _o_geop(ix, g, a):
    from oracle import dyn_core as O

    O.compute_geopotential(ix, a["zh"], a["gz"])


register(StageSpec("compute_geopotential", "ComputeGeopotential#0", ("gz",), _o_geop,
                   lambda sf, qf, rt, q, d: rt.call("fv3_compute_geopotential", q["zh"].ptr, q["gz"].ptr)))


def _o_fxadv(ix, g, a):
    from oracle import fxadv as O

    O.fv_prep(ix, g, a["uc"], a["vc"], a["crx"], a["cry"], a["x_area_flux"], a["y_area_flux"], a["uc_contra"],
              a["vc_contra"], float(a["dt"]))


def _n_fxadv(sf, qf, rt, q, d):
    rt.call("fv3_fv_prep", q["uc"].ptr, q["vc"].ptr, q["crx"].ptr, q["cry"].ptr, q["x_area_flux"].ptr,
            q["y_area_flux"].ptr, q["uc_contra"].ptr, q["vc_contra"].ptr, f(d, "dt"))


# uc_contra / vc_contra are compared where they are consumed downstream (the reference leaves stale data
# from the previous call in a few halo points next to tile edges, fxadv.py:33-48)
register(StageSpec("fv_prep", "FxAdv#0", ("crx", "cry", "x_area_flux", "y_area_flux", "uc_contra", "vc_contra"),
                   _o_fxadv, _n_fxadv,
                   regions={"uc_contra": (slice(3, 3 + NX + 1), slice(None)), "vc_contra": (slice(None), slice(3, 3 + NX + 1))}))


# ---------------------------------------------------------------------------------------------
# fvtp2d: golden call n -> (hord, nord column, damp column, levels); order of calls in d_sw.py:935-1237,
# updatedzd.py:283-356 and tracer_2d_1l.py:264-392
FVTP2D_CASES = {0: (6, "nord_v", "damp_vt", NZ), 1: (6, None, None, NZ), 2: (6, "nord_t", "damp_t", NZ),
                3: (6, "nord_v", "damp_vt", NZ), 4: (6, None, None, NZ), 5: (6, None, None, NZ + 1),
                6: (8, None, None, NZ)}
FLUX_REGIONS = {"q_x_flux": (slice(3, 3 + NX + 1), slice(3, 3 + NX)), "q_y_flux": (slice(3, 3 + NX), slice(3, 3 + NX + 1))}


def _columns():
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.stencils.d_sw import get_column_namelist

    return get_column_namelist(baroclinic_config(NX).d_grid_shallow_water, NZ)


def _make_fvtp2d(n):
    hord, no, da, nk = FVTP2D_CASES[n]

    def oracle(ix, g, a):
        from oracle import fvtp2d as O

        col = _columns()
        O.fvtp2d(ix, g, a["q"], a["crx"], a["cry"], a["x_area_flux"], a["y_area_flux"], a["q_x_flux"], a["q_y_flux"],
                 hord, nk, a.get("x_mass_flux"), a.get("y_mass_flux"), a.get("mass"), col[no] if no else None,
                 col[da] if da else None, float(g["damp_da_min"]))

    def native(sf, qf, rt, q, d):
        from pace_b200.fv3core.stencils.fvtp2d import FiniteVolumeTransport

        col = _columns()
        tp = FiniteVolumeTransport(sf, qf, rt.grid_data, rt.damping, 0, hord, col[no] if no else None,
                                   col[da] if da else None)
        tp(q["q"], q["crx"], q["cry"], q["x_area_flux"], q["y_area_flux"], q["q_x_flux"], q["q_y_flux"],
           q.get("x_mass_flux"), q.get("y_mass_flux"), q.get("mass"), nk=nk)

    register(StageSpec(f"fvtp2d_{n}", f"FvTp2d#{n}", ("q_x_flux", "q_y_flux"), oracle, native, regions=FLUX_REGIONS,
                       check_untouched=False))


for _n in FVTP2D_CASES:
    _make_fvtp2d(_n)


def _make_delnflux(n, nord_name):
    def oracle(ix, g, a):
        from oracle import fvtp2d as O

        O.delnflux_nosg(ix, g, a["q"], a["fx2"], a["fy2"], a["damp_c"], _columns()[nord_name], NZ, d2=a["d2"])

    def native(sf, qf, rt, q, d):
        import torch

        from pace_b200.fv3core.stencils.fvtp2d import DelnFluxNoSG, _column

        dn = DelnFluxNoSG(sf, rt.damping, rt.grid_data.rarea, _columns()[nord_name])
        dn(q["q"], q["fx2"], q["fy2"], _column(rt, d["in.damp_c"][:NZ]))

    register(StageSpec(f"delnflux_nosg_{n}", f"DelnFluxNoSG#{n}", ("fx2", "fy2"), oracle, native,
                       regions={"fx2": (slice(3, 3 + NX + 1), slice(3, 3 + NX)), "fy2": (slice(3, 3 + NX), slice(3, 3 + NX + 1))},
                       check_untouched=False))


# call order within one acoustic substep: #0 inside fvtp2d_dp (nord_v), #1 delnflux_nosg_w, #4 delnflux_nosg_v
_make_delnflux(0, "nord_v")
_make_delnflux(1, "nord_w")
_make_delnflux(4, "nord_v")


# ---------------------------------------------------------------------------------------------
CORNERS = (slice(3, 3 + NX + 1), slice(3, 3 + NX + 1))
COMPUTE = (slice(3, 3 + NX), slice(3, 3 + NX))


def _dsw_cols(rt):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.stencils.d_sw import ColumnNamelist

    return ColumnNamelist(rt, baroclinic_config(NX).d_grid_shallow_water, rt.damping)


def _n_a2b(kstart):
    def native(sf, qf, rt, q, d):
        rt.call("fv3_a2b_ord4", q["qin"].ptr, q["qout"].ptr, kstart, NZ - kstart)

    return native


def _o_todo(ix, g, a):
    import pytest

    pytest.skip("numpy oracle for this stage not written yet (native kernels are pinned on the reference golden)")


register(StageSpec("a2b_ord4_0", "A2B_Ord4#0", ("qout",), _o_todo, _n_a2b(3), regions={"qout": CORNERS}, tol=1e-13))


def _n_divdamp(sf, qf, rt, q, d):
    cols = _dsw_cols(rt)
    rt.call("fv3_divergence_damping", q["u"].ptr, q["v"].ptr, q["va"].ptr, q["damped_rel_vort_bgrid"].ptr, q["ua"].ptr,
            q["divg_d"].ptr, q["vc"].ptr, q["uc"].ptr, q["delpc"].ptr, q["ke"].ptr, q["rel_vort_agrid"].ptr, f(d, "dt"),
            cols.ref)


def _o_divdamp(ix, g, a):
    from oracle import divergence_damping as O
    from pace_b200.fv3core._config import baroclinic_config

    col = _columns()
    cfg = baroclinic_config(NX).d_grid_shallow_water
    nord_col = np.asarray(col["nord"])
    k0 = int(np.argmax(nord_col > 0)) if (nord_col > 0).any() else NZ
    O.divergence_damping(ix, g, a["u"], a["v"], a["va"], a["damped_rel_vort_bgrid"], a["ua"], a["divg_d"], a["vc"], a["uc"],
                         a["delpc"], a["ke"], a["rel_vort_agrid"], float(a["dt"]), np.asarray(col["d2_divg"]), k0,
                         int(nord_col.max()), cfg.dddmp, cfg.d4_bg)


register(StageSpec("divergence_damping", "DivergenceDamping#0", ("damped_rel_vort_bgrid", "ke", "delpc", "divg_d"), _o_divdamp,
                   _n_divdamp, regions={n: CORNERS for n in ("damped_rel_vort_bgrid", "ke", "delpc", "divg_d")}, tol=1e-13,
                   check_untouched=False))


def _n_dsw(sf, qf, rt, q, d):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.stencils.d_sw import DGridShallowWaterLagrangianDynamics

    cfg = baroclinic_config(NX).d_grid_shallow_water
    dsw = DGridShallowWaterLagrangianDynamics(sf, qf, rt.grid_data, rt.damping, _dsw_cols(rt), False, False, cfg)
    dsw(*[q[n] for n in ("delpc", "delp", "pt", "u", "v", "w", "uc", "vc", "ua", "va", "divgd", "mfx", "mfy", "cx", "cy",
                         "crx", "cry", "xfx", "yfx", "q_con", "zh", "heat_source", "diss_est")], f(d, "dt"))


_XI = (slice(3, 3 + NX + 1), slice(3, 3 + NX))      # x-interface fields on the compute rows
_YI = (slice(3, 3 + NX), slice(3, 3 + NX + 1))
def _o_dsw(ix, g, a):
    from oracle import d_sw as O
    from pace_b200.fv3core._config import baroclinic_config

    O.d_sw(ix, g, a, _columns(), baroclinic_config(NX).d_grid_shallow_water)


register(StageSpec(
    "d_sw", "D_SW#0",
    ("delp", "pt", "w", "q_con", "u", "v", "mfx", "mfy", "cx", "cy", "crx", "cry", "xfx", "yfx", "heat_source", "diss_est"),
    _o_dsw, _n_dsw, tol=1e-12, check_untouched=False,
    regions={"delp": COMPUTE, "pt": COMPUTE, "w": COMPUTE, "q_con": COMPUTE, "u": _YI, "v": _XI, "mfx": _XI, "mfy": _YI,
             "cx": _XI, "cy": _YI, "crx": (slice(3, 3 + NX + 1), slice(None)), "cry": (slice(None), slice(3, 3 + NX + 1)),
             "xfx": (slice(3, 3 + NX + 1), slice(None)), "yfx": (slice(None), slice(3, 3 + NX + 1)),
             "heat_source": COMPUTE, "diss_est": COMPUTE}))


def _n_dzd(sf, qf, rt, q, d):
    from pace_b200.fv3core.stencils.updatedzd import UpdateHeightOnDGrid

    up = UpdateHeightOnDGrid(sf, qf, rt.damping, rt.grid_data, 0, 6, _dsw_cols(rt))
    up(q["surface_height"], q["height"], q["courant_number_x"], q["courant_number_y"], q["x_area_flux"], q["y_area_flux"],
       q["ws"], f(d, "dt"))


def _o_dzd(ix, g, a):
    from oracle import updatedz as O

    col = _columns()
    O.update_dz_d(ix, g, a["surface_height"], a["height"], a["courant_number_x"], a["courant_number_y"], a["x_area_flux"],
                  a["y_area_flux"], a["ws"], float(a["dt"]), col["damp_vt"], col["nord_v"], 6)


register(StageSpec("update_dz_d", "UpdateDzD#0", ("height", "ws"), _o_dzd, _n_dzd, tol=1e-13,
                   regions={"height": COMPUTE, "ws": COMPUTE}, check_untouched=False))


# ---------------------------------------------------------------------------------------------
# remaining acoustic-substep stages; second-step goldens (non-trivial w, q_con, heat_source ...)
S2 = "c12s2"
RING1 = None


def _n_riem3(sf, qf, rt, q, d):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.stencils.acoustic_misc import NonhydrostaticVerticalSolver

    solver = NonhydrostaticVerticalSolver(sf, qf, baroclinic_config(NX).riemann)
    solver(bool(d["in.last_call"]), f(d, "dt"), q["cappa"], f(d, "ptop"), q["zs"], q["ws"], q["delz"], q["q_con"], q["delp"],
           q["pt"], q["zh"], q["p"], q["ppe"], q["pk3"], q["pk"], q["log_p_interface"], q["w"])


def _o_riem3(ix, g, a):
    from oracle import riem_solver as O

    O.riem_solver3(bool(a["last_call"]), float(a["dt"]), a["cappa"], float(a["ptop"]), a["zs"], a["ws"], a["delz"], a["q_con"],
                   a["delp"], a["pt"], a["zh"], a["p"], a["ppe"], a["pk3"], a["pk"], a["log_p_interface"], a["w"], 0.05,
                   ix.nx, ix.ny, ix.nz)


register(StageSpec("riem_solver3", "Riem_Solver3#0", ("delz", "zh", "w", "p", "ppe", "pk3", "pk", "log_p_interface"), _o_riem3,
                   _n_riem3, tol=1e-11, near_zero=1e-12, check_untouched=False, case=S2,
                   # w is a small difference of large terms: the reference itself allows 5e-6 for this routine on every
                   # backend (fv3core/tests/savepoint/translate/overrides/standard.yaml, Riem_Solver3)
                   tols={"w": 5e-6},
                   regions={n: COMPUTE for n in ("delz", "zh", "w", "p", "ppe", "pk3", "pk", "log_p_interface")}))
def _o_edge_pe(ix, g, a):
    from oracle import dyn_core as O

    O.edge_pe(ix, a["pe"], a["delp"], float(a["ptop"]))


def _o_pk3_halo(ix, g, a):
    from oracle import dyn_core as O

    O.pk3_halo(ix, a["pk3"], a["delp"], float(a["ptop"]), float(a["akap"]))


def _o_diffusive_heating(ix, g, a):
    from oracle import dyn_core as O

    O.apply_diffusive_heating(ix, a["delp"], a["delz"], a["cappa"], a["heat_source"], a["pt"], float(a["delt_time_factor"]))


register(StageSpec("edge_pe", "PE_Halo#0", ("pe",), _o_edge_pe,
                   lambda sf, qf, rt, q, d: rt.call("fv3_edge_pe", q["pe"].ptr, q["delp"].ptr, f(d, "ptop")), case=S2))
register(StageSpec("pk3_halo", "PK3_Halo#0", ("pk3",), _o_pk3_halo,
                   lambda sf, qf, rt, q, d: rt.call("fv3_pk3_halo", q["pk3"].ptr, q["delp"].ptr, f(d, "ptop"), f(d, "akap")),
                   tol=1e-13, case=S2))
def _o_nhpg(ix, g, a):
    from oracle import a2b as O

    O.nh_p_grad(ix, g, a["u"], a["v"], a["pp"], a["gz"], a["pk3"], a["delp"], float(a["dt"]), float(a["ptop"]), float(a["akap"]))


register(StageSpec("nh_p_grad", "NH_P_Grad#0", ("u", "v", "pp", "gz", "pk3"), _o_nhpg,
                   lambda sf, qf, rt, q, d: rt.call("fv3_nh_p_grad", q["u"].ptr, q["v"].ptr, q["pp"].ptr, q["gz"].ptr,
                                                    q["pk3"].ptr, q["delp"].ptr, f(d, "dt"), f(d, "ptop"), f(d, "akap")),
                   tol=1e-11, near_zero=1e-12, case=S2, check_untouched=False,
                   regions={"u": _YI, "v": _XI, "pp": CORNERS, "gz": CORNERS, "pk3": CORNERS}))


def _n_ray(sf, qf, rt, q, d):
    from pace_b200.fv3core.stencils.acoustic_misc import RayleighDamping

    RayleighDamping(sf, 3000.0, 10.0, False)(q["u"], q["v"], q["w"], None, None, f(d, "dt"), f(d, "ptop"))


def _o_ray(ix, g, a):
    from oracle import dyn_core as O

    O.ray_fast(ix, a["u"], a["v"], a["w"], a["dp"], a["pfull"], float(a["dt"]), float(a["ptop"]), 3000.0, 10.0)


register(StageSpec("ray_fast", "Ray_Fast#0", ("u", "v", "w"), _o_ray, _n_ray, tol=1e-13, case=S2,
                   regions={"u": _YI, "v": _XI, "w": COMPUTE}, check_untouched=False))


def _n_del2(nmax):
    def native(sf, qf, rt, q, d):
        rt.call("fv3_del2cubed", q["qdel"].ptr, f(d, "cd"), nmax, NZ)

    return native


def _o_del2(nmax):
    def oracle(ix, g, a):
        from oracle import dyn_core as O

        O.del2cubed(ix, g, a["qdel"], float(a["cd"]), nmax)

    return oracle


register(StageSpec("del2cubed_heat", "Del2Cubed#0", ("qdel",), _o_del2(3), _n_del2(3), tol=1e-13, case=S2,
                   regions={"qdel": COMPUTE}))
register(StageSpec("del2cubed_omga", "Del2Cubed#1", ("qdel",), _o_del2(1), _n_del2(1), tol=1e-13, case=S2,
                   regions={"qdel": COMPUTE}))
register(StageSpec("diffusive_heating", "DiffusiveHeating#0", ("pt",), _o_diffusive_heating,
                   lambda sf, qf, rt, q, d: rt.call("fv3_apply_diffusive_heating", q["delp"].ptr, q["delz"].ptr, q["cappa"].ptr,
                                                    q["heat_source"].ptr, q["pt"].ptr, f(d, "delt_time_factor"), NZ),
                   tol=1e-13, case=S2, regions={"pt": COMPUTE}))


# Every spec defined on the first step is run on the SECOND step of the same reference run (non-trivial w, q_con,
# heat source ...); only a few keep their first-step variant too, to bound the size of the committed fixtures.
KEEP_FIRST_STEP = {"riem_solver_c", "c_sw"}
DROP = {"fvtp2d_0", "fvtp2d_2", "fvtp2d_4", "delnflux_nosg_0", "delnflux_nosg_1", "a2b_ord4_0"}
for _name, _spec in list(SPECS.items()):
    if _spec.case == "c12":
        if _name not in DROP:
            SPECS[_name + "@s2"] = dataclasses.replace(_spec, name=_name + "@s2", case=S2)
        if _name not in KEEP_FIRST_STEP:
            del SPECS[_name]


def golden_stage_files():
    """{case: sorted stage file names} needed by the registered specs (used by tests/golden/make_committed.py)."""
    out = {}
    for sp in SPECS.values():
        out.setdefault(sp.case, set()).add(sp.golden)
    return {k: sorted(v) for k, v in out.items()}


# ---------------------------------------------------------------------------------------------
# vertical remapping (second-step goldens)
def _ptr_array(rt, qs):
    import torch

    return torch.tensor([q.ptr for q in qs], dtype=torch.int64).to(rt.device)


def _make_map_single(n, iv, i_extra=0, j_extra=0):
    def native(sf, qf, rt, q, d):
        qs = q.get("qs")
        rt.call("fv3_map_single", q["q1"].ptr, q["pe1"].ptr, q["pe2"].ptr, qs.ptr if (qs is not None and iv == -2) else None,
                1, float(d.get("in.qmin", 0.0)), 9, iv, i_extra, j_extra)

    def oracle(ix, g, a):
        from oracle import remap as O

        O.map_single(a["q1"], a["pe1"], a["pe2"], a.get("qs") if iv == -2 else None, float(a.get("qmin", 0.0)), iv,
                     ix.nx, ix.ny, ix.nz, i_extra, j_extra)

    reg = (slice(3, 3 + NX + i_extra), slice(3, 3 + NX + j_extra))
    SPECS[f"map_single_{n}"] = StageSpec(f"map_single_{n}", f"MapSingle#{n}", ("q1",), oracle, native, tol=1e-12, near_zero=1e-13,
                                         case=S2, regions={"q1": reg}, check_untouched=False)


_make_map_single(0, 1)        # pt
_make_map_single(1, 0)        # qvapor
_make_map_single(9, -2)       # w
_make_map_single(10, 1)       # delz
_make_map_single(11, -1, j_extra=1)   # u
_make_map_single(12, -1, i_extra=1)   # v


def _n_fillz(sf, qf, rt, q, d):
    names = ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke"]
    arr = _ptr_array(rt, [q["tracers." + n] for n in names])
    rt.call("fv3_fillz", arr.data_ptr(), 8, q["dp2"].ptr)


def _o_fillz(ix, g, a):
    from oracle import remap as O

    for n in ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke"]:
        O.fillz(a["tracers." + n], a["dp2"], ix.nx, ix.ny, ix.nz)


SPECS["fillz"] = StageSpec("fillz", "Fillz#0", tuple("tracers." + n for n in ("qvapor", "qliquid", "qrain", "qice")), _o_fillz,
                           _n_fillz, tol=5e-6, case=S2, regions={"tracers." + n: COMPUTE for n in ("qvapor", "qliquid", "qrain", "qice")},
                           check_untouched=False)


def _n_fvsetup(sf, qf, rt, q, d):
    arr = _ptr_array(rt, [q[n] for n in ("qvapor", "qliquid", "qrain", "qsnow", "qice", "qgraupel")])
    rt.call("fv3_fv_setup", arr.data_ptr(), q["q_con"].ptr, q["cvm"].ptr, q["pkz"].ptr, q["pt"].ptr, q["cappa"].ptr,
            q["delp"].ptr, q["delz"].ptr, q["dp1"].ptr)


def _fvsetup_expected(d):
    """fv3_fv_setup also applies pt_to_potential_density_pt: expected pt comes from the PtAdjust golden."""
    d2 = H.load_stage(S2, 0, "PtAdjust#0")
    d = dict(d)
    d["out.pt"] = d2["out.pt"]
    return d


def _o_fvsetup(ix, g, a):
    """The reference's fv_setup alone (pt is only adjusted by the next reference stencil, which the native call fuses)."""
    from oracle import remap as O

    O.fv_setup(a["qvapor"], a["qliquid"], a["qrain"], a["qsnow"], a["qice"], a["qgraupel"], a["q_con"], a["cvm"], a["pkz"],
               a["pt"], a["cappa"], a["delp"], a["delz"], a["dp1"], ix.nx, ix.ny, ix.nz)


SPECS["fv_setup"] = StageSpec("fv_setup", "FVSetup#0", ("q_con", "cvm", "pkz", "cappa", "dp1", "pt"), _o_fvsetup, _n_fvsetup, tol=1e-13,
                              case=S2, regions={n: COMPUTE for n in ("q_con", "cvm", "pkz", "cappa", "dp1", "pt")},
                              check_untouched=False, expected=_fvsetup_expected)


def _n_remap(sf, qf, rt, q, d):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.stencils.remapping import LagrangianToEulerian

    tr = {n: q["tracers." + n] for n in ("qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke")}
    l2e = LagrangianToEulerian(sf, qf, baroclinic_config(NX).remapping, rt.grid_data.area_64, 8, None, tr)
    l2e(tr, q["pt"], q["delp"], q["delz"], q["peln"], q["u"], q["v"], q["w"], q["cappa"], q["q_con"], q["q_cld"], q["pkz"],
        q["pk"], q["pe"], q["hs"], q["ps"], q["wsd"], None, None, q["dp1"], f(d, "ptop"), f(d, "akap"), f(d, "zvir"),
        bool(d["in.last_step"]), f(d, "consv_te"), f(d, "mdt"))


def _o_remap(ix, g, a):
    from oracle import remap as O

    O.lagrangian_to_eulerian(a, ix.nx, ix.ny, ix.nz)


_RM_OUT = ("pt", "delp", "delz", "peln", "u", "v", "w", "cappa", "q_con", "pkz", "pk", "pe", "ps", "tracers.qvapor")
SPECS["remapping"] = StageSpec(
    "remapping", "Remapping#0", _RM_OUT, _o_remap, _n_remap, tol=1e-11, near_zero=1e-13, case=S2, check_untouched=False,
    tols={"w": 5e-6},
    regions={**{n: COMPUTE for n in _RM_OUT}, "u": _YI, "v": _XI})
