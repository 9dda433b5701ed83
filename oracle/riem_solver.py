"""Oracle: semi-implicit nonhydrostatic column solvers (test infrastructure).

Follows fv3core/pace/fv3core/stencils/sim1_solver.py:20-141, riem_solver_c.py:21-123,172-250 and
riem_solver3.py:26-145,207-321 of the reference.
"""
import numpy as np

from .constants import GRAV, RDGAS


def sim1_solver(w, dm, gm, dz, pt, pm, pem, ws, cp3, dt, p_fac):
    """sim1_solver.py:20-141.  Arrays [ni, nj, nz(+1)]; w, dz updated in place; returns pe [ni, nj, nz+1]."""
    nz = dm.shape[2]
    t1g = 2.0 * dt * dt
    rdt = 1.0 / dt
    sh = dm.shape[:2]
    pe = np.zeros(sh + (nz + 1,))
    pe[:, :, :nz] = np.exp(gm * np.log(-dm / dz * RDGAS * pt)) - pm            # :55-60
    w1 = w.copy()
    g_rat = np.zeros(sh + (nz,))
    bb = np.zeros(sh + (nz,))
    dd = np.zeros(sh + (nz,))
    g_rat[:, :, : nz - 1] = dm[:, :, : nz - 1] / dm[:, :, 1:nz]                 # :61-68
    bb[:, :, : nz - 1] = 2.0 * (1.0 + g_rat[:, :, : nz - 1])
    dd[:, :, : nz - 1] = 3.0 * (pe[:, :, : nz - 1] + g_rat[:, :, : nz - 1] * pe[:, :, 1:nz])
    bb[:, :, nz - 1] = 2.0
    dd[:, :, nz - 1] = 3.0 * pe[:, :, nz - 1]
    pp = np.zeros(sh + (nz + 1,))
    gam = np.zeros(sh + (nz,))
    aa = np.zeros(sh + (nz,))
    bet = bb[:, :, 0].copy()                                                     # :70-92
    pp[:, :, 1] = dd[:, :, 0] / bet
    bets = np.zeros(sh + (nz,))
    bets[:, :, 0] = bet
    for k in range(1, nz):
        gam[:, :, k] = g_rat[:, :, k - 1] / bets[:, :, k - 1]
        bets[:, :, k] = bb[:, :, k] - gam[:, :, k]
    for k in range(2, nz + 1):
        pp[:, :, k] = (dd[:, :, k - 1] - pp[:, :, k - 1]) / bets[:, :, k - 1]
    for k in range(nz - 1, 0, -1):
        pp[:, :, k] = pp[:, :, k] - gam[:, :, k] * pp[:, :, k + 1]
        aa[:, :, k] = t1g * 0.5 * (gm[:, :, k - 1] + gm[:, :, k]) / (dz[:, :, k - 1] + dz[:, :, k]) * (pem[:, :, k] + pp[:, :, k])
    bet = dm[:, :, 0] - aa[:, :, 1]                                              # :95-122
    w[:, :, 0] = (dm[:, :, 0] * w1[:, :, 0] + dt * pp[:, :, 1]) / bet
    for k in range(1, nz - 1):
        gam[:, :, k] = aa[:, :, k] / bet
        bet = dm[:, :, k] - (aa[:, :, k] + aa[:, :, k + 1] + aa[:, :, k] * gam[:, :, k])
        w[:, :, k] = (dm[:, :, k] * w1[:, :, k] + dt * (pp[:, :, k + 1] - pp[:, :, k]) - aa[:, :, k] * w[:, :, k - 1]) / bet
    k = nz - 1
    p1 = t1g * gm[:, :, k] / dz[:, :, k] * (pem[:, :, k + 1] + pp[:, :, k + 1])
    gam[:, :, k] = aa[:, :, k] / bet
    bet = dm[:, :, k] - (aa[:, :, k] + p1 + aa[:, :, k] * gam[:, :, k])
    w[:, :, k] = (dm[:, :, k] * w1[:, :, k] + dt * (pp[:, :, k + 1] - pp[:, :, k]) - p1 * ws - aa[:, :, k] * w[:, :, k - 1]) / bet
    for k in range(nz - 2, -1, -1):
        w[:, :, k] = w[:, :, k] - gam[:, :, k + 1] * w[:, :, k + 1]
    pe[:, :, 0] = 0.0                                                            # :123-130
    for k in range(1, nz + 1):
        pe[:, :, k] = pe[:, :, k - 1] + dm[:, :, k - 1] * (w[:, :, k - 1] - w1[:, :, k - 1]) * rdt
    p1 = (pe[:, :, nz - 1] + 2.0 * pe[:, :, nz]) * 1.0 / 3.0                     # :131-137
    p1s = np.zeros(sh + (nz,))
    p1s[:, :, nz - 1] = p1
    for k in range(nz - 2, -1, -1):
        p1s[:, :, k] = (pe[:, :, k] + bb[:, :, k] * pe[:, :, k + 1] + g_rat[:, :, k] * pe[:, :, k + 2]) * 1.0 / 3.0 - g_rat[:, :, k] * p1s[:, :, k + 1]
    maxp = np.where(p_fac * dm > p1s + pm, p_fac * pm, p1s + pm)                  # :138-145
    dz[...] = -dm * RDGAS * pt * np.exp((cp3 - 1.0) * np.log(maxp))
    return pe


def riem_solver_c(dt2, cappa, ptop, hs, ws, ptc, q_con, delpc, gz, pef, w3, p_fac, nx, ny, nz, halo=3):
    """NonhydrostaticVerticalSolverCGrid.__call__ (riem_solver_c.py:172-250); gz, pef updated in place
    over the compute domain + 1 halo cell."""
    si = slice(halo - 1, halo + nx + 1)
    sj = slice(halo - 1, halo + ny + 1)
    dm = delpc[si, sj, :nz].copy()                                               # precompute :56-87
    w = w3[si, sj, :nz].copy()
    pem = np.zeros(dm.shape[:2] + (nz + 1,))
    peg = np.zeros_like(pem)
    pem[:, :, 0] = ptop
    peg[:, :, 0] = ptop
    for k in range(1, nz + 1):
        pem[:, :, k] = pem[:, :, k - 1] + dm[:, :, k - 1]
        peg[:, :, k] = peg[:, :, k - 1] + dm[:, :, k - 1] * (1.0 - q_con[si, sj, k - 1])
    dz = gz[si, sj, 1 : nz + 1] - gz[si, sj, :nz]
    gm = 1.0 / (1.0 - cappa[si, sj, :nz])
    dm = dm / GRAV
    pm = (peg[:, :, 1:] - peg[:, :, :nz]) / np.log(peg[:, :, 1:] / peg[:, :, :nz])
    pe = sim1_solver(w, dm, gm, dz, ptc[si, sj, :nz], pm, pem, ws[si, sj], cappa[si, sj, :nz], dt2, p_fac)
    pef[si, sj, 0] = ptop                                                         # finalize :113-123
    pef[si, sj, 1 : nz + 1] = pe[:, :, 1:] + pem[:, :, 1:]
    gz[si, sj, nz] = hs[si, sj]
    for k in range(nz - 1, -1, -1):
        gz[si, sj, k] = gz[si, sj, k + 1] - dz[:, :, k] * GRAV


def riem_solver3(last_call, dt, cappa, ptop, zs, ws, delz, q_con, delp, pt, zh, pe, ppe, pk3, pk, peln, w, p_fac, nx, ny, nz,
                 halo=3):
    """NonhydrostaticVerticalSolver.__call__ (riem_solver3.py:207-321): precompute (:26-90), Sim1Solver, finalize
    (:93-145); compute domain; delz, zh, w, ppe, pk3 (and pe, pk, peln on the last call) updated in place."""
    from .constants import KAPPA

    si = slice(halo, halo + nx)
    sj = slice(halo, halo + ny)
    dm = delp[si, sj, :nz].copy()
    sh = dm.shape[:2]
    pem = np.zeros(sh + (nz + 1,))
    peg = np.zeros_like(pem)
    pem[:, :, 0] = ptop
    peg[:, :, 0] = ptop
    for k in range(1, nz + 1):
        pem[:, :, k] = pem[:, :, k - 1] + dm[:, :, k - 1]
        peg[:, :, k] = peg[:, :, k - 1] + dm[:, :, k - 1] * (1.0 - q_con[si, sj, k - 1])
    peln1 = np.log(ptop)
    pelng = np.log(peg)
    pelng[:, :, 0] = peln1
    lp = np.log(pem)
    lp[:, :, 0] = peln1
    pk3v = np.exp(KAPPA * lp)
    pk3v[:, :, 0] = np.exp(KAPPA * peln1)
    pk3[si, sj, : nz + 1] = pk3v
    if last_call:
        peln[si, sj, : nz + 1] = lp
        pk[si, sj, : nz + 1] = pk3v
        pe[si, sj, : nz + 1] = pem
    pm = (peg[:, :, 1:] - peg[:, :, :nz]) / (pelng[:, :, 1:] - pelng[:, :, :nz])
    dz = zh[si, sj, 1 : nz + 1] - zh[si, sj, :nz]
    gm = 1.0 / (1.0 - cappa[si, sj, :nz])
    dm = dm * (1.0 / GRAV)
    wv = w[si, sj, :nz].copy()
    pp = sim1_solver(wv, dm, gm, dz, pt[si, sj, :nz], pm, pem, ws[si, sj], cappa[si, sj, :nz], dt, p_fac)
    w[si, sj, :nz] = wv
    delz[si, sj, :nz] = dz
    ppe[si, sj, : nz + 1] = pp
    zh[si, sj, nz] = zs[si, sj]
    for k in range(nz - 1, -1, -1):
        zh[si, sj, k] = zh[si, sj, k + 1] - dz[:, :, k]
