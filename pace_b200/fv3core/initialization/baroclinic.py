"""Analytic Jablonowski-Williamson baroclinic-wave initial state on the cubed sphere.

Same call interface and result as the reference's `init_baroclinic_state`
(fv3core/pace/fv3core/initialization/baroclinic.py:405-539, formulas of
baroclinic_jablonowski_williamson.py:19-167), restated as batched numpy over all local subdomains (init-time host
code; the state is then uploaded and its halos filled by the device halo exchange).
Validated against the reference's own state in tests/test_grid_generation.py.
"""
import math

import numpy as np

from ... import constants as c
from ...util.grid import generation as gen
from ..dycore_state import FIELDS, DycoreState

NH = c.N_HALO_DEFAULT
U0 = 35.0
PCEN = (math.pi / 9.0, 2.0 * math.pi / 9.0)
U1 = 1.0
ETA_0 = 0.252
ETA_SURFACE = 1.0
ETA_TROPOPAUSE = 0.2
T_0 = 288.0
DELTA_T = 480000.0
LAPSE_RATE = 0.005
SURFACE_PRESSURE = 1.0e5
R_PERT = c.RADIUS / 10.0


def _zonal_wind(eta_v, lon, lat):
    """Eq. (2) + Gaussian perturbation, eq. (10) of JW2006; lon/lat [s, i, j] -> [s, i, j, k]."""
    u = U0 * np.cos(eta_v) ** (3.0 / 2.0) * np.sin(2.0 * lat[..., None]) ** 2.0
    r = gen.gc_dist(PCEN[0], lon, PCEN[1], lat)[..., None]
    near = np.broadcast_to((r / R_PERT) ** 2.0 < 40.0, u.shape)
    pert = np.broadcast_to(U1 * np.exp(-((r / R_PERT) ** 2.0)), u.shape)
    return np.where(near, u + pert, u)


def _project(u, lon, e):
    """Zonal wind -> component along grid unit vector e (baroclinic.py:57-69)."""
    return u * (e[..., 1] * np.cos(lon) - e[..., 0] * np.sin(lon))[..., None]


def _t_mean(eta):
    t = T_0 * eta ** (c.RDGAS * LAPSE_RATE / c.GRAV)
    top = ETA_TROPOPAUSE > eta
    t[top] = t[top] + DELTA_T * (ETA_TROPOPAUSE - eta[top]) ** 5.0
    return t


def _lat_terms(lat):
    a = -2.0 * (np.sin(lat) ** 6.0) * (np.cos(lat) ** 2.0 + 1.0 / 3.0) + 10.0 / 63.0
    b = (8.0 / 5.0) * (np.cos(lat) ** 3.0) * (np.sin(lat) ** 2.0 + 2.0 / 3.0) - math.pi / 4.0
    return a, b


def _temperature(eta, eta_v, t_mean, lat):
    a, b = _lat_terms(lat[..., None])
    return t_mean + 0.75 * (eta * math.pi * U0 / c.RDGAS) * np.sin(eta_v) * np.sqrt(np.cos(eta_v)) * (
        a * 2.0 * U0 * np.cos(eta_v) ** (3.0 / 2.0) + b * c.RADIUS * c.OMEGA)


def _surface_geopotential(lat):
    eta_vs = (ETA_SURFACE - ETA_0) * math.pi * 0.5
    uc = U0 * (np.cos(eta_vs) ** (3.0 / 2.0))
    a, b = _lat_terms(lat)
    return uc * (a * uc + b * c.RADIUS * c.OMEGA)


def _nine_point(fn, lon, lat, lat_a):
    """Weighted 9-point cell average of fn(lat) (baroclinic.py:153-213); lon/lat corners [s, n+1, n+1]."""
    lat2 = gen.midpoint_ll(lon[:, :-1, :], lon[:, 1:, :], lat[:, :-1, :], lat[:, 1:, :])[1]
    lat3 = gen.midpoint_ll(lon[:, 1:, :-1], lon[:, 1:, 1:], lat[:, 1:, :-1], lat[:, 1:, 1:])[1]
    lat4 = gen.midpoint_ll(lon[:, :-1, 1:], lon[:, 1:, 1:], lat[:, :-1, 1:], lat[:, 1:, 1:])[1]
    lat5 = gen.midpoint_ll(lon[:, :, :-1], lon[:, :, 1:], lat[:, :, :-1], lat[:, :, 1:])[1]
    p1 = fn(lat_a)
    p2, p3, p4, p5 = fn(lat2[:, :, :-1]), fn(lat3), fn(lat4), fn(lat5[:, :-1, :])
    p6, p7, p8, p9 = fn(lat[:, :-1, :-1]), fn(lat[:, 1:, :-1]), fn(lat[:, 1:, 1:]), fn(lat[:, :-1, 1:])
    return 0.25 * p1 + 0.125 * (p2 + p3 + p4 + p5) + 0.0625 * (p6 + p7 + p8 + p9)


def baroclinic_arrays(grids, adiabatic=False, hydrostatic=False, moist_phys=True):
    """Initial state arrays (reference storage order [s, i, j(, k)], halos NOT exchanged) for the given per-subdomain
    metric-term dicts (pace_b200.util.grid.generation.generate / the reference's GridData dump)."""
    if hydrostatic:
        raise NotImplementedError("the hot path is non-hydrostatic")
    S = len(grids)
    ni, nj = grids[0]["lon"].shape
    nx, ny = ni - 2 * NH - 1, nj - 2 * NH - 1
    ak, bk = np.asarray(grids[0]["ak"], dtype=np.float64), np.asarray(grids[0]["bk"], dtype=np.float64)
    nk = len(ak)
    ptop = float(ak[0])
    shape3, shape2 = (S, ni, nj, nk), (S, ni, nj)
    st = {}
    for name, (dims, _) in FIELDS.items():
        st[name] = np.zeros(shape3 if len(dims) == 3 else shape2)
    st["delp"][:] = 1e30
    for a, b in ((slice(None, NH), slice(None, NH)), (slice(None, NH), slice(NH + ny, None)),
                 (slice(NH + nx, None), slice(None, NH)), (slice(NH + nx, None), slice(NH + ny, None))):
        st["delp"][:, a, b] = 0.0
    st["pt"][:] = 1.0
    st["ua"][:] = 1e35
    st["va"][:] = 1e35
    st["uc"][:] = 1e30
    st["vc"][:] = 1e30
    st["w"][:] = 1.0e30
    st["delz"][:] = 1.0e25
    st["phis"][:] = 1.0e25
    st["ps"][:] = SURFACE_PRESSURE
    ci, cj = slice(NH, NH + nx), slice(NH, NH + ny)
    bi, bj = slice(NH, NH + nx + 1), slice(NH, NH + ny + 1)
    # vertical columns (horizontally uniform surface pressure)
    delp_col = np.full(nk, 1e30)
    delp_col[:-1] = ak[1:] - ak[:-1] + SURFACE_PRESSURE * (bk[1:] - bk[:-1])
    pe = np.empty(nk)
    pe[0] = ptop
    for k in range(1, nk):
        pe[k] = pe[k - 1] + delp_col[k - 1]
    peln = np.empty(nk)
    peln[0] = math.log(ptop)
    peln[1:] = np.log(pe[1:])
    pk = np.empty(nk)
    pk[0] = ptop ** c.KAPPA
    pk[1:] = np.exp(c.KAPPA * np.log(pe[1:]))
    eta = np.zeros(nk)
    eta_v = np.zeros(nk)
    eta[:-1] = 0.5 * ((ak[:-1] + ak[1:]) / SURFACE_PRESSURE + bk[:-1] + bk[1:])
    eta_v[:-1] = (eta[:-1] - ETA_0) * math.pi * 0.5
    st["delp"][:, ci, cj, :-1] = delp_col[:-1]
    st["pe"][:, ci, cj, :] = pe
    st["peln"][:, ci, cj, :] = peln
    st["pk"][:, ci, cj, :] = pk
    dpeln = peln[1:] - peln[:-1]

    stack = lambda k, si, sj: np.stack([np.asarray(g[k])[si, sj] for g in grids])  # noqa: E731
    lon, lat = stack("lon", bi, bj), stack("lat", bi, bj)
    lat_a = stack("lat_agrid", ci, cj)
    ee1, ee2, es1, ew2 = (stack(k, bi, bj) for k in ("ee1", "ee2", "es1", "ew2"))
    # D-grid winds: Simpson-like average of the projected zonal wind at the two end points and the edge midpoint
    a = slice(0, nx + 1)
    mlon, mlat = gen.midpoint_ll(lon[:, :, :-1], lon[:, :, 1:], lat[:, :, :-1], lat[:, :, 1:])
    uu1 = _project(_zonal_wind(eta_v, lon[:, a, 1:], lat[:, a, 1:]), lon[:, a, 1:], ee2[:, a, 1:])
    uu3 = _project(_zonal_wind(eta_v, lon[:, a, :-1], lat[:, a, :-1]), lon[:, a, :-1], ee2[:, a, :-1])
    uu2 = _project(_zonal_wind(eta_v, mlon, mlat), mlon, ew2[:, a, :ny])
    st["v"][:, bi, cj, :] = 0.25 * (uu1 + 2.0 * uu2 + uu3)
    mlon, mlat = gen.midpoint_ll(lon[:, :-1, :], lon[:, 1:, :], lat[:, :-1, :], lat[:, 1:, :])
    uu1 = _project(_zonal_wind(eta_v, lon[:, :-1, :], lat[:, :-1, :]), lon[:, :-1, :], ee1[:, :-1, :])
    uu3 = _project(_zonal_wind(eta_v, lon[:, 1:, :], lat[:, 1:, :]), lon[:, 1:, :], ee1[:, 1:, :])
    uu2 = _project(_zonal_wind(eta_v, mlon, mlat), mlon, es1[:, :nx, :])
    st["u"][:, ci, bj, :] = 0.25 * (uu1 + 2.0 * uu2 + uu3)
    # temperature and surface geopotential: 9-point cell means
    tm = _t_mean(eta)
    pt = _nine_point(lambda la: _temperature(eta, eta_v, tm, la), lon, lat, lat_a)
    st["phis"][:, ci, cj] = _nine_point(_surface_geopotential, lon, lat, lat_a)
    st["w"][:, ci, cj, :] = 0.0
    if not adiabatic:
        ptmp = delp_col[:-1] / dpeln - SURFACE_PRESSURE
        q = 0.021 * np.exp(-((lat_a[..., None] / PCEN[1]) ** 4.0)) * np.exp(-((ptmp / 34000.0) ** 2.0))
        st["qvapor"][:, ci, cj, :-1] = q
        pt = pt / (1.0 + c.ZVIR * st["qvapor"][:, ci, cj, :])
    st["pt"][:, ci, cj, :] = pt
    # p_var (baroclinic.py:331-364)
    st["ps"][:, ci, cj] = pe[-1]
    delz = c.RDG * pt[..., :-1] * dpeln
    st["delz"][:, ci, cj, :-1] = delz
    if moist_phys:
        arg = c.RDG * delp_col[:-1] * pt[..., :-1] * (1.0 + c.ZVIR * st["qvapor"][:, ci, cj, :-1]) / delz
    else:
        arg = c.RDG * delp_col[:-1] * pt[..., :-1] / delz
    st["pkz"][:, ci, cj, :-1] = np.exp(c.KAPPA * np.log(arg))
    return st


def fill_tracers(state_arrays, scale=0.1):
    """Benchmark config 4: make all 8 advected tracers non-trivial, q_m = qvapor*(m+1)/10."""
    from ..dycore_state import TRACER_VARIABLES

    for m, name in enumerate(TRACER_VARIABLES[1:8], start=1):
        state_arrays[name][:] = state_arrays["qvapor"] * (m + 1) * scale
    return state_arrays


def init_baroclinic_state(grid_data, quantity_factory, adiabatic: bool, hydrostatic: bool, moist_phys: bool, comm,
                          fill_all_tracers: bool = False) -> DycoreState:
    """DycoreState of all local subdomains; phis and (u, v) halos exchanged as in the reference (:533-537)."""
    grids = grid_data.host_dicts()
    arrays = baroclinic_arrays(grids, adiabatic, hydrostatic, moist_phys)
    if fill_all_tracers:
        fill_tracers(arrays)
    per_rank = [{k: v[s] for k, v in arrays.items()} for s in range(len(grids))]
    state = DycoreState.init_from_numpy_arrays(per_rank, quantity_factory)
    comm.halo_update(state.phis, NH)
    comm.vector_halo_update(state.u, state.v, NH)
    return state
