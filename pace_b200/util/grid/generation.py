"""Cubed-sphere grid generation: gnomonic equal-distance grid + every metric term the hot path reads.

Product-side replacement of the reference's `MetricTerms` (util/pace/util/grid/generation.py:200-2366, with
gnomonic.py, mirror.py, geometry.py) for the terms in `GridData` / `DampingCoefficients`.  Design differences:

  * everything is computed ONCE PER CUBE TILE (6 arrays of (N+7)^2 points, halo 3) with vectorised numpy, then cut
    into the 6*layout^2 subdomains; the reference computes per rank with halo exchanges between ranks.  A halo cell
    that lies inside a tile is the neighbour subdomain's own value in both schemes, so the cut is exact.
  * halo fills across tile edges use the gather tables of pace_b200.util.topology (rotation folded into indices).

Init-time host code (fp64 numpy), not on the hot path.  Validated against the reference's own arrays for
c12 layout (1,1) and c24 layout (2,2) in tests/test_grid_generation.py.
"""
import os
from typing import Dict, List

import numpy as np

from ... import constants as c
from .. import topology

PI = c.PI
RADIUS = c.RADIUS
H = c.N_HALO_DEFAULT
BIG = 1.0e8
TINY = 1.0e-8


# ------------------------------------------------------------------------------------------------------------
# spherical geometry helpers (gnomonic.py:167-262, 329-373, 593-702)

def _unit_scale(v):
    """v * (1/|v|) — the multiply-by-reciprocal normalisation of gnomonic.normalize_vector."""
    return v * (1.0 / np.sqrt((v ** 2.0).sum(-1, keepdims=True)))


def _unit_div(v):
    """v / |v| — gnomonic.normalize_xyz."""
    return v / np.sqrt((v ** 2).sum(-1, keepdims=True))


def ll2xyz(lon, lat):
    lon, lat = np.broadcast_arrays(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64))
    v = np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)
    return _unit_scale(v)


def xyz2ll(xyz):
    p = _unit_div(np.asarray(xyz))
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    lon = np.where(np.abs(x) + np.abs(y) >= 1.0e-10, np.arctan2(y, x), 0.0)
    lon = np.where(lon < 0.0, lon + 2 * PI, lon)
    return lon, np.arcsin(z)


def midpoint_xyz(*pts):
    return _unit_div(sum(pts))


def midpoint_ll(lon1, lon2, lat1, lat2):
    return xyz2ll(midpoint_xyz(ll2xyz(lon1, lat1), ll2xyz(lon2, lat2)))


def gc_beta(lon1, lon2, lat1, lat2):
    return np.arcsin(np.sqrt(np.sin((lat1 - lat2) / 2.0) ** 2
                             + np.cos(lat1) * np.cos(lat2) * np.sin((lon1 - lon2) / 2.0) ** 2)) * 2.0


def gc_dist(lon1, lon2, lat1, lat2, radius=RADIUS):
    return radius * gc_beta(lon1, lon2, lat1, lat2)


def gc_dist_xyz(p1, p2, radius=RADIUS):
    lon1, lat1 = xyz2ll(p1)
    lon2, lat2 = xyz2ll(p2)
    return gc_dist(lon1, lon2, lat1, lat2, radius)


def _sph_cos(pc, p2, p3):
    p = np.cross(pc, p2)
    q = np.cross(pc, p3)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (p * q).sum(-1) / np.sqrt((p ** 2).sum(-1) * (q ** 2).sum(-1))


def _sph_angle(pc, p2, p3):
    with np.errstate(invalid="ignore"):
        a = np.arccos(_sph_cos(pc, p2, p3))
    return np.where(np.isnan(a), 0.0, a)


def rect_area(p1, p2, p3, p4, radius=RADIUS):
    tot = _sph_angle(p2, p3, p1)
    for q1, q2, q3 in ((p3, p2, p4), (p4, p3, p1), (p1, p4, p2)):
        tot = tot + _sph_angle(q1, q2, q3)
    return (tot - 2 * PI) * radius ** 2


def tri_area(p1, p2, p3, radius=RADIUS):
    tot = _sph_angle(p1, p2, p3)
    for q1, q2, q3 in ((p2, p3, p1), (p3, p1, p2)):
        tot = tot + _sph_angle(q1, q2, q3)
    return (tot - PI) * radius ** 2


# ------------------------------------------------------------------------------------------------------------
# gnomonic equal-distance face + symmetrisation + rotation to the 6 tiles
# (gnomonic.local_gnomonic_ed :26-153, mirror.mirror_grid :9-214, generation._init_dgrid :1476-1619)

def _face_lonlat(N):
    alpha = np.arcsin(3 ** -0.5)
    rsq3 = 3 ** -0.5
    dely = np.multiply(2.0, alpha / float(N))
    lat_w = -alpha + dely * np.arange(N + 1, dtype=np.float64)
    lon_w, lon_e = 0.75 * PI, 1.25 * PI
    lat_s, lat_n = lat_w[0], -alpha + dely * float(N)

    def project(p):  # central projection on the cube face x = -1/sqrt(3)
        return -p[..., 1] * rsq3 / p[..., 0], -p[..., 2] * rsq3 / p[..., 0]

    pw = ll2xyz(lon_w, lat_w)
    y_w, z_w = project(pw)
    # south edge: mirror image of the west edge in the plane through the SW and NE corners and the centre
    p1, p2 = ll2xyz(lon_w, lat_s), ll2xyz(lon_e, lat_n)
    nb = np.cross(p1, p2)
    nb = nb / np.sqrt((nb ** 2).sum())
    pm = pw - 2.0 * (pw * nb).sum(-1, keepdims=True) * nb
    lon_s, lat_se = xyz2ll(pm)
    y_s, z_s = project(ll2xyz(lon_s, lat_se))
    sw, se = ll2xyz(lon_w, lat_s), ll2xyz(lon_e, lat_s)
    y_s[0], z_s[0] = sw[1], sw[2]
    y_s[N], z_s[N] = se[1], se[2]
    pp = np.empty((N + 1, N + 1, 3))
    pp[..., 0] = -rsq3
    pp[..., 1] = y_s[:, None]
    pp[..., 2] = z_w[None, :]
    pp[0, :, 1] = y_w
    pp[:, 0, 2] = z_s
    pp[0, 0, 1:] = sw[1:]
    lon, lat = xyz2ll(pp)
    return lon - PI, lat


def _symmetrise(lon, lat, N):
    def avg(a):
        m = 0.25 * (np.abs(a) + np.abs(a[::-1, :]) + np.abs(a[:, ::-1]) + np.abs(a[::-1, ::-1]))
        return np.copysign(m, a)

    lon, lat = avg(lon), avg(lat)
    if N % 2 == 0:
        lon[N // 2, :] = 0.0
    return lon, lat


def _rot(axis, lon, lat, angle_deg):
    """mirror._rot_3d with the reference's left-handed spherical convention (z = -r sin(lat))."""
    r = RADIUS + 0.0 * lon
    x, y, z = r * np.cos(lon) * np.cos(lat), r * np.sin(lon) * np.cos(lat), -r * np.sin(lat)
    a = np.deg2rad(angle_deg)
    co, si = np.cos(a), np.sin(a)
    if axis == 1:
        x2, y2, z2 = x, co * y + si * z, -si * y + co * z
    elif axis == 2:
        x2, y2, z2 = co * x - si * z, y, si * x + co * z
    else:
        x2, y2, z2 = co * x + si * y, -si * x + co * y, z
    rr = np.sqrt(x2 * x2 + y2 * y2 + z2 * z2)
    lon2 = np.where(np.abs(x2) + np.abs(y2) < 1.0e-10, 0.0, np.arctan2(y2, x2))
    return lon2, np.arccos(z2 / rr) - PI / 2.0


def tile_corner_lonlat(N):
    """(lon, lat)[6, N+1, N+1] of the cell corners of the six tiles (radians)."""
    lon0, lat0 = _symmetrise(*_face_lonlat(N), N)
    lons, lats = [lon0.copy()], [lat0.copy()]
    seqs = {1: [(3, -90.0)], 2: [(3, -90.0), (1, 90.0)], 3: [(3, -180.0), (1, 90.0)], 4: [(3, 90.0), (2, 90.0)],
            5: [(2, 90.0), (3, 0.0)]}
    m = N // 2
    for t in range(1, 6):
        lo, la = lon0, lat0
        for ax, ang in seqs[t]:
            lo, la = _rot(ax, lo, la, ang)
        lo, la = lo.copy(), la.copy()
        if N % 2 == 0:  # force pole / dateline consistency (mirror.py:113-131,152-156)
            if t == 2:
                la[m, m] = PI / 2.0
                lo[: m + 1, m] = 0.0
                lo[m + 1:, m] = PI
            elif t == 3:
                lo[:, m] = PI
        lons.append(lo)
        lats.append(la)
    lon, lat = np.stack(lons), np.stack(lats)
    lon = lon - PI / 18.0   # shift the corner away from Japan (generation.py:1606-1611)
    lon = np.where(lon < 0, lon + 2 * PI, lon)
    lon[np.abs(lon) < 1e-10] = 0.0
    lat[np.abs(lat) < 1e-10] = 0.0
    return lon, lat


# ------------------------------------------------------------------------------------------------------------
# tile-level halo fills

class _TileHalo:
    def __init__(self, N):
        self.dec = topology.Decomposition(N, 1)
        self._t = {}

    def _table(self, sx, sy):
        k = (sx, sy)
        if k not in self._t:
            self._t[k] = topology.build_halo_table(self.dec, H, sx, sy, halo=H)
        return self._t[k]

    def scalar(self, a, stagger):
        topology.apply_table_numpy(self._table(stagger, None), [a[t] for t in range(6)])

    def vector(self, ax, ay, sx, sy):
        topology.apply_table_numpy(self._table(sx, sy), [ax[t] for t in range(6)], [ay[t] for t in range(6)])


def _corner_loop():
    return [(i, j) for i in range(1, H + 1) for j in range(1, H + 1)]


def _fill_corners_b_x(q, N):
    """corners.fill_corners_2d_bgrid, direction x (corners.py:717-760) on [.., i, j, ...] arrays (axis 1, 2 of q)."""
    isc, jsc, iec, jec = H, H, H + N - 1, H + N - 1
    for i, j in _corner_loop():
        q[:, isc - i, jsc - j] = q[:, isc - j, jsc + i]
        q[:, isc - i, jec + 1 + j] = q[:, isc - j, jec + 1 - i]
        q[:, iec + 1 + i, jsc - j] = q[:, iec + 1 + j, jsc + i]
        q[:, iec + 1 + i, jec + 1 + j] = q[:, iec + 1 + j, jec + 1 - i]


def _fill_corners_a(q, N, direction):
    """corners.fill_corners_2d_agrid (corners.py:763-812)."""
    isc, jsc, iec, jec = H, H, H + N - 1, H + N - 1
    for i, j in _corner_loop():
        if direction == "x":
            q[:, isc - i, jsc - j] = q[:, isc - j, jsc + i - 1]
            q[:, isc - i, jec + j] = q[:, isc - j, jec - i + 1]
            q[:, iec + i, jsc - j] = q[:, iec + j, isc + i - 1]
            q[:, iec + i, jec + j] = q[:, iec + j, jec - i + 1]
        else:
            q[:, isc - j, jsc - i] = q[:, isc + i - 1, jsc - j]
            q[:, isc - j, jec + i] = q[:, isc + i - 1, jec + j]
            q[:, iec + j, jsc - i] = q[:, iec - i + 1, jsc - j]
            q[:, iec + j, jec + i] = q[:, iec - i + 1, jec + j]


def _fill_corners_agrid_pair(x, y, N):
    """corners.fill_corners_agrid with vector=False (corners.py:849-885)."""
    n = H
    ie, je = H + N - 1, H + N - 1
    for i, j in _corner_loop():
        x[:, n - i, n - j] = y[:, n - j, n - 1 + i]
        y[:, n - j, n - i] = x[:, n - 1 + i, n - j]
        x[:, n - i, je + j] = y[:, n - j, je - i + 1]
        y[:, n - j, je + i] = x[:, n - 1 + i, je + j]
        x[:, ie + i, n - j] = y[:, ie + j, n - 1 + i]
        y[:, ie + j, n - i] = x[:, ie - i + 1, n - j]
        x[:, ie + i, je + j] = y[:, ie + j, je - i + 1]
        y[:, ie + j, je + i] = x[:, ie - i + 1, je + j]


def _fill_corners_dgrid_pair(x, y, N):
    """corners.fill_corners_dgrid with vector=False (corners.py:888-932)."""
    isc, jsc, iec, jec = H, H, H + N - 1, H + N - 1
    for i, j in _corner_loop():
        x[:, isc - i, jsc - j] = y[:, isc - j, i + 2]
        y[:, isc - i, jsc - j] = x[:, j + 2, jsc - i]
        x[:, isc - i, jec + 1 + j] = y[:, isc - j, jec + 1 - i]
        y[:, isc - i, jec + j] = x[:, j + 2, jec + 1 + i]
        x[:, iec + i, jsc - j] = y[:, iec + 1 + j, i + 2]
        y[:, iec + 1 + i, jsc - j] = x[:, iec - j + 1, jsc - i]
        x[:, iec + i, jec + 1 + j] = y[:, iec + 1 + j, jec - i + 1]
        y[:, iec + 1 + i, jec + j] = x[:, iec - j + 1, jec + 1 + i]


def _fill_corners_cgrid_pair(x, y, N):
    """corners.fill_corners_cgrid with vector=False (corners.py:935-978)."""
    isc, jsc, iec, jec = H, H, H + N - 1, H + N - 1
    for i, j in _corner_loop():
        x[:, isc - i, jsc - j] = y[:, j + 2, jsc - i]
        y[:, isc - i, jsc - j] = x[:, isc - j, i + 2]
        x[:, isc - i, jec + j] = y[:, j + 2, jec + 1 + i]
        y[:, isc - i, jec + 1 + j] = x[:, isc - j, jec + 1 - i]
        x[:, iec + 1 + i, jsc - j] = y[:, iec + 1 - j, jsc - i]
        y[:, iec + i, jsc - j] = x[:, iec + 1 + j, i + 2]
        x[:, iec + 1 + i, jec + j] = y[:, iec + 1 - j, jec + 1 + i]
        y[:, iec + i, jec + 1 + j] = x[:, iec + 1 + j, jec + 1 - i]


def _set_halo_corners(a, value):
    """geometry._fill_halo_corners on [6, i, j, ...] arrays whose cell extent is a.shape[1] x a.shape[2]."""
    a[:, :H, :H] = value
    a[:, :H, -H:] = value
    a[:, -H:, :H] = value
    a[:, -H:, -H:] = value


# ------------------------------------------------------------------------------------------------------------

def load_eta(nz):
    """ak, bk of the hybrid pressure coordinate (util/pace/util/grid/eta.py:24-573); 79 levels only."""
    if nz != 79:
        raise NotImplementedError("only the 79-level vertical grid is tabulated")
    t = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "eta_l79.txt"))
    return t[:, 0].copy(), t[:, 1].copy()


def generate_tiles(N: int) -> Dict[str, np.ndarray]:
    """All horizontal metric terms on the six tiles: name -> [6, N+7, N+7(, 3)] in reference storage order [i, j]."""
    ni = N + 2 * H + 1
    halo = _TileHalo(N)
    out: Dict[str, np.ndarray] = {}
    A, B = (1, 1), (0, 0)
    XD, YD = (1, 0), (0, 1)   # stagger of (X_DIM, Y_INTERFACE_DIM) and (X_INTERFACE_DIM, Y_DIM) fields

    # --- D-grid (corner) positions, generation._init_dgrid
    grid = np.zeros((6, ni, ni, 2))
    glon, glat = tile_corner_lonlat(N)
    grid[:, H:H + N + 1, H:H + N + 1, 0] = glon
    grid[:, H:H + N + 1, H:H + N + 1, 1] = glat
    halo.scalar(grid, B)
    _fill_corners_b_x(grid, N)
    lon, lat = grid[..., 0], grid[..., 1]
    # --- A-grid (centre) positions, generation._init_agrid
    dxyz = ll2xyz(lon, lat)
    agrid = np.zeros((6, ni, ni, 2))
    cen = midpoint_xyz(dxyz[:, 1:, 1:], dxyz[:, :-1, :-1], dxyz[:, 1:, :-1], dxyz[:, :-1, 1:])
    agrid[:, :-1, :-1, 0], agrid[:, :-1, :-1, 1] = xyz2ll(cen)
    halo.scalar(agrid, A)
    _fill_corners_a(agrid[..., 0], N, "x")
    _fill_corners_a(agrid[..., 1], N, "y")
    lon_a, lat_a = agrid[..., 0], agrid[..., 1]
    axyz = ll2xyz(lon_a[:, :-1, :-1], lat_a[:, :-1, :-1])
    out.update(lon=lon, lat=lat, lon_agrid=lon_a, lat_agrid=lat_a)

    # --- dx, dy (generation._compute_dxdy)
    cs = slice(H, H + N + 1)
    dx = np.zeros((6, ni, ni))
    dy = np.zeros((6, ni, ni))
    dx[:, H:H + N, cs] = gc_dist(lon[:, H:H + N, cs], lon[:, H + 1:H + N + 1, cs], lat[:, H:H + N, cs], lat[:, H + 1:H + N + 1, cs])
    dy[:, cs, H:H + N] = gc_dist(lon[:, cs, H:H + N], lon[:, cs, H + 1:H + N + 1], lat[:, cs, H:H + N], lat[:, cs, H + 1:H + N + 1])
    halo.vector(dx, dy, XD, YD)
    np.abs(dx, out=dx)
    np.abs(dy, out=dy)
    _fill_corners_dgrid_pair(dx, dy, N)
    # --- dxa, dya (generation._compute_dxdy_agrid)
    lyc, tyc = midpoint_ll(lon[:, :, :-1], lon[:, :, 1:], lat[:, :, :-1], lat[:, :, 1:])
    dxa_t = gc_dist(lyc[:, :-1], lyc[:, 1:], tyc[:, :-1], tyc[:, 1:])
    lxc, txc = midpoint_ll(lon[:, :-1, :], lon[:, 1:, :], lat[:, :-1, :], lat[:, 1:, :])
    dya_t = gc_dist(lxc[:, :, :-1], lxc[:, :, 1:], txc[:, :, :-1], txc[:, :, 1:])
    _fill_corners_agrid_pair(dxa_t, dya_t, N)
    dxa = np.zeros((6, ni, ni))
    dya = np.zeros((6, ni, ni))
    dxa[:, :-1, :-1], dya[:, :-1, :-1] = dxa_t, dya_t
    halo.vector(dxa, dya, A, A)
    np.abs(dxa, out=dxa)
    np.abs(dya, out=dya)
    # --- dxc, dyc (generation._compute_dxdy_center)
    la, ta = lon_a[:, :-1, :-1], lat_a[:, :-1, :-1]
    dxc_t = gc_dist(la[:, :-1], la[:, 1:], ta[:, :-1], ta[:, 1:])
    dyc_t = gc_dist(la[:, :, :-1], la[:, :, 1:], ta[:, :, :-1], ta[:, :, 1:])
    dxc = np.zeros((6, ni, ni))
    dyc = np.zeros((6, ni, ni))
    dxc[:, 1:-1, :-1] = dxc_t
    dxc[:, 0, :-1] = dxc_t[:, 0]
    dxc[:, -1, :-1] = dxc_t[:, -1]
    dyc[:, :-1, 1:-1] = dyc_t
    dyc[:, :-1, 0] = dyc_t[:, :, 0]
    dyc[:, :-1, -1] = dyc_t[:, :, -1]
    ce = slice(H, H + N)      # compute cells
    # tile-edge values: twice the distance from the edge midpoint to the first cell centre (gnomonic.py:547-590)
    for side, idx_d, idx_a in (("w", H, H), ("e", H + N, H + N - 1)):
        edge_pt = 0.5 * (dxyz[:, idx_d, H + 1:H + N + 1] + dxyz[:, idx_d, H:H + N])
        dxc[:, idx_d, ce] = 2 * gc_dist_xyz(edge_pt, axyz[:, idx_a, ce])
    for side, idx_d, idx_a in (("s", H, H), ("n", H + N, H + N - 1)):
        edge_pt = 0.5 * (dxyz[:, H + 1:H + N + 1, idx_d] + dxyz[:, H:H + N, idx_d])
        dyc[:, ce, idx_d] = 2 * gc_dist_xyz(edge_pt, axyz[:, ce, idx_a])
    halo.vector(dxc, dyc, YD, XD)
    np.abs(dxc, out=dxc)
    np.abs(dyc, out=dyc)
    _fill_corners_cgrid_pair(dxc, dyc, N)
    out.update(dx=dx, dy=dy, dxa=dxa, dya=dya, dxc=dxc, dyc=dyc)

    # --- areas (generation._compute_area, _compute_area_c)
    area = np.full((6, ni, ni), -1.0e8)
    P = dxyz[:, H:H + N + 1, H:H + N + 1]
    area[:, ce, ce] = rect_area(P[:, :-1, :-1], P[:, :-1, 1:], P[:, 1:, 1:], P[:, 1:, :-1])
    halo.scalar(area, A)
    area_c = np.zeros((6, ni, ni))
    Q = ll2xyz(lon_a[:, 2:-3, 2:-3], lat_a[:, 2:-3, 2:-3])     # centres H-1 .. H+N
    ac = rect_area(Q[:, :-1, :-1], Q[:, :-1, 1:], Q[:, 1:, 1:], Q[:, 1:, :-1])
    ll_, lr_, ul_, ur_ = Q[:, :-1, :-1], Q[:, 1:, :-1], Q[:, :-1, 1:], Q[:, 1:, 1:]
    ac[:, 0, 0] = tri_area(ul_[:, 0, 0], ur_[:, 0, 0], lr_[:, 0, 0])
    ac[:, -1, 0] = tri_area(ur_[:, -1, 0], ul_[:, -1, 0], ll_[:, -1, 0])
    ac[:, -1, -1] = tri_area(lr_[:, -1, -1], ll_[:, -1, -1], ul_[:, -1, -1])
    ac[:, 0, -1] = tri_area(ll_[:, 0, -1], lr_[:, 0, -1], ur_[:, 0, -1])
    # tile borders: twice the half cell on this side of the edge (gnomonic.py:419-510), in the reference's order W,N,E,S
    D2 = dxyz[:, 2:-2, 2:-2]
    A2 = axyz[:, 2:-2, 2:-2]

    def west_edge(Dg, Ag):
        yc = 0.5 * (Dg[:, 1, :-1] + Dg[:, 1, 1:])
        return 2 * rect_area(yc[:, :-1], Ag[:, 1, :-1], Ag[:, 1, 1:], yc[:, 1:])

    ac[:, 0, :] = west_edge(D2, A2)
    ac[:, :, -1] = west_edge(D2[:, :, ::-1].transpose(0, 2, 1, 3), A2[:, :, ::-1].transpose(0, 2, 1, 3))
    ac[:, -1, :] = west_edge(D2[:, ::-1], A2[:, ::-1])
    ac[:, :, 0] = west_edge(D2.transpose(0, 2, 1, 3), A2.transpose(0, 2, 1, 3))
    area_c[:, H:-H, H:-H] = ac
    halo.scalar(area_c, B)
    _fill_corners_b_x(area_c, N)
    out.update(area=area, area_64=area, area_c=area_c)

    # --- unit vectors at cell centres (geometry.get_center_vector :14-60)
    cpts = midpoint_xyz(dxyz[:, :-1, :-1], dxyz[:, 1:, :-1], dxyz[:, :-1, 1:], dxyz[:, 1:, 1:])
    p1 = midpoint_xyz(dxyz[:, :-1, :-1], dxyz[:, :-1, 1:])
    p2 = midpoint_xyz(dxyz[:, 1:, :-1], dxyz[:, 1:, 1:])
    ec1 = _unit_div(np.cross(cpts, np.cross(p2, p1)))
    p1 = midpoint_xyz(dxyz[:, :-1, :-1], dxyz[:, 1:, :-1])
    p2 = midpoint_xyz(dxyz[:, :-1, 1:], dxyz[:, 1:, 1:])
    ec2 = _unit_div(np.cross(cpts, np.cross(p2, p1)))
    _set_halo_corners(ec1, BIG)
    _set_halo_corners(ec2, BIG)

    # --- edge unit vectors ew (x-interfaces) and es (y-interfaces) (geometry.py:63-146)
    pp = midpoint_xyz(dxyz[:, 1:-1, :-1], dxyz[:, 1:-1, 1:])
    p2 = np.cross(axyz[:, :-1], axyz[:, 1:])
    p2[:, H - 1] = np.cross(pp[:, H - 1], axyz[:, H])
    p2[:, -H] = np.cross(axyz[:, -H - 1], pp[:, -H])
    ew1 = np.zeros((6, ni, ni - 1, 3))
    ew2 = np.zeros((6, ni, ni - 1, 3))
    ew1[:, 1:-1] = _unit_div(np.cross(p2, pp))
    ew2[:, 1:-1] = _unit_div(np.cross(np.cross(dxyz[:, 1:-1, :-1], dxyz[:, 1:-1, 1:]), pp))
    pp = midpoint_xyz(dxyz[:, :-1, 1:-1], dxyz[:, 1:, 1:-1])
    p2 = np.cross(axyz[:, :, :-1], axyz[:, :, 1:])
    p2[:, :, H - 1] = np.cross(pp[:, :, H - 1], axyz[:, :, H])
    p2[:, :, -H] = np.cross(axyz[:, :, -H - 1], pp[:, :, -H])
    es1 = np.zeros((6, ni - 1, ni, 3))
    es2 = np.zeros((6, ni - 1, ni, 3))
    es2[:, :, 1:-1] = _unit_div(np.cross(p2, pp))
    es1[:, :, 1:-1] = _unit_div(np.cross(np.cross(dxyz[:, :-1, 1:-1], dxyz[:, 1:, 1:-1]), pp))
    for a in (ew1, ew2, es1, es2):
        _set_halo_corners(a, 0.0)
    full = np.full((6, ni, ni, 3), np.nan)
    out["ew2"] = full.copy()
    out["ew2"][:, 1:-1, :-1] = ew2[:, 1:-1]
    out["es1"] = full.copy()
    out["es1"][:, :-1, 1:-1] = es1[:, :, 1:-1]
    # --- corner unit vectors ee1, ee2 (geometry.calculate_xy_unit_vectors :267-309)
    cx = np.cross(dxyz[:, H - 1:-H - 1, H:-H], dxyz[:, H + 1:ni - H + 1, H:-H])
    cx[:, 0] = np.cross(dxyz[:, H, H:-H], dxyz[:, H + 1, H:-H])
    cx[:, -1] = np.cross(dxyz[:, -H - 2, H:-H], dxyz[:, -H - 1, H:-H])
    cy = np.cross(dxyz[:, H:-H, H - 1:-H - 1], dxyz[:, H:-H, H + 1:ni - H + 1])
    cy[:, :, 0] = np.cross(dxyz[:, H:-H, H], dxyz[:, H:-H, H + 1])
    cy[:, :, -1] = np.cross(dxyz[:, H:-H, -H - 2], dxyz[:, H:-H, -H - 1])
    out["ee1"] = full.copy()
    out["ee2"] = full.copy()
    out["ee1"][:, H:-H, H:-H] = _unit_div(np.cross(cx, dxyz[:, H:-H, H:-H]))
    out["ee2"][:, H:-H, H:-H] = _unit_div(np.cross(cy, dxyz[:, H:-H, H:-H]))

    # --- supergrid angles (geometry.calculate_supergrid_cos_sin :149-236); last axis = points 1..9
    d00, d10, d01, d11 = dxyz[:, :-1, :-1], dxyz[:, 1:, :-1], dxyz[:, :-1, 1:], dxyz[:, 1:, 1:]
    cos_sg = np.zeros((6, ni - 1, ni - 1, 9))
    cos_sg[..., 5] = _sph_cos(d00, d10, d01)
    cos_sg[..., 6] = -1 * _sph_cos(d10, d00, d11)
    cos_sg[..., 7] = _sph_cos(d11, d10, d01)
    cos_sg[..., 8] = -1 * _sph_cos(d01, d00, d11)
    cos_sg[..., 0] = _sph_cos(midpoint_xyz(d00, d01), axyz, d01)
    cos_sg[..., 1] = _sph_cos(midpoint_xyz(d00, d10), d10, axyz)
    cos_sg[..., 2] = _sph_cos(midpoint_xyz(d10, d11), axyz, d10)
    cos_sg[..., 3] = _sph_cos(midpoint_xyz(d01, d11), d01, axyz)
    cos_sg[..., 4] = (ec1 * ec2).sum(-1)
    with np.errstate(invalid="ignore"):
        cos_sg[np.abs(1.0 - cos_sg) < 1e-15] = 1.0
        s2 = 1.0 - cos_sg ** 2
        s2[s2 < 0] = 0.0
        sin_sg = np.sqrt(s2)
        sin_sg[sin_sg > 1.0] = 1.0
    n = H
    sin_sg[:, n - 1, :n, 2] = sin_sg[:, :n, n, 1]
    sin_sg[:, :n, n - 1, 3] = sin_sg[:, n, :n, 0]
    sin_sg[:, n - 1, -n:, 2] = sin_sg[:, :n, -n - 1, 3][:, ::-1]
    sin_sg[:, :n, -n, 1] = sin_sg[:, n, -n - 2:ni - 1 - n + 1, 0]
    sin_sg[:, -n, :n, 0] = sin_sg[:, -n:, n, 1][:, ::-1]
    sin_sg[:, -n:, n - 1, 3] = sin_sg[:, -n - 1, :n, 2][:, ::-1]
    sin_sg[:, -n, -n:, 0] = sin_sg[:, -n:, -n - 1, 3]
    sin_sg[:, -n:, -n, 1] = sin_sg[:, -n - 1, -n:, 2]

    # --- derived trig terms (geometry.calculate_trig_uv :312-413)
    cosa = np.full((6, ni, ni), BIG)
    sina = np.full((6, ni, ni), BIG)
    cosa[:, n:-n, n:-n] = 0.5 * (cos_sg[:, n - 1:-n, n - 1:-n, 7] + cos_sg[:, n:ni - 1 - n + 1, n:ni - 1 - n + 1, 5])
    sina[:, n:-n, n:-n] = 0.5 * (sin_sg[:, n - 1:-n, n - 1:-n, 7] + sin_sg[:, n:ni - 1 - n + 1, n:ni - 1 - n + 1, 5])
    cosa_u = np.full((6, ni, ni - 1), BIG)
    sina_u = np.full((6, ni, ni - 1), BIG)
    rsin_u = np.full((6, ni, ni - 1), BIG)
    cosa_v = np.full((6, ni - 1, ni), BIG)
    sina_v = np.full((6, ni - 1, ni), BIG)
    rsin_v = np.full((6, ni - 1, ni), BIG)
    cosa_u[:, 1:-1] = 0.5 * (cos_sg[:, :-1, :, 2] + cos_sg[:, 1:, :, 0])
    sina_u[:, 1:-1] = 0.5 * (sin_sg[:, :-1, :, 2] + sin_sg[:, 1:, :, 0])
    rsin_u[:, 1:-1] = 1.0 / np.maximum(sina_u[:, 1:-1] ** 2, TINY)
    cosa_v[:, :, 1:-1] = 0.5 * (cos_sg[:, :, :-1, 3] + cos_sg[:, :, 1:, 1])
    sina_v[:, :, 1:-1] = 0.5 * (sin_sg[:, :, :-1, 3] + sin_sg[:, :, 1:, 1])
    rsin_v[:, :, 1:-1] = 1.0 / np.maximum(sina_v[:, :, 1:-1] ** 2, TINY)
    cosa_s = cos_sg[..., 4].copy()
    rsin2 = 1.0 / np.maximum(sin_sg[..., 4] ** 2, TINY)
    _set_halo_corners(cosa_s, BIG)
    rsina = 1.0 / np.maximum(sina[:, n:-n, n:-n] ** 2, TINY)

    def lim(a):
        a = a.copy()
        small = np.abs(a) < TINY
        a[small] = TINY * np.sign(a[small])
        return a

    rsina[:, 0, :] = BIG
    rsin_u[:, n] = 1.0 / lim(sina_u[:, n])
    rsina[:, -1, :] = BIG
    rsin_u[:, -n - 1] = 1.0 / lim(sina_u[:, -n - 1])
    rsina[:, :, 0] = BIG
    rsin_v[:, :, n] = 1.0 / lim(sina_v[:, :, n])
    rsina[:, :, -1] = BIG
    rsin_v[:, :, -n - 1] = 1.0 / lim(sina_v[:, :, -n - 1])

    def pad(a, value=0.0):
        f = np.full((6, ni, ni), value)
        f[:, :a.shape[1], :a.shape[2]] = a
        return f

    rs = np.zeros((6, ni, ni))
    rs[:, n:-n, n:-n] = rsina
    out.update(cosa=cosa, sina=sina, cosa_u=pad(cosa_u), cosa_v=pad(cosa_v), cosa_s=pad(cosa_s), sina_u=pad(sina_u),
               sina_v=pad(sina_v), rsin_u=pad(rsin_u), rsin_v=pad(rsin_v), rsina=rs, rsin2=pad(rsin2))

    # --- geometry.supergrid_corner_fix :416-497 (after the derived terms, as generation._init_cell_trigonometry)
    for sg, val in ((sin_sg, TINY), (cos_sg, BIG)):
        _set_halo_corners(sg, val)
        # SW
        sg[:, n - 1, :n, 2] = sg[:, :n, n, 1]
        sg[:, :n, n - 1, 3] = sg[:, n, :n, 0]
        # NW (j mirrored)
        v = sg[:, :, ::-1]
        v[:, :n, n - 1, 1] = v[:, n, :n, 0]
        v[:, n - 1, :n, 2] = v[:, :n, n, 3]
        # SE (i mirrored)
        v = sg[:, ::-1]
        v[:, n - 1, :n, 0] = v[:, :n, n, 1]
        v[:, :n, n - 1, 3] = v[:, n, :n, 2]
        # NE (both mirrored)
        v = sg[:, ::-1, ::-1]
        v[:, n - 1, :n, 0] = v[:, :n, n, 3]
        v[:, :n, n - 1, 1] = v[:, n, :n, 2]
    for k in range(9):
        out[f"cos_sg{k + 1}"] = pad(cos_sg[..., k])
        out[f"sin_sg{k + 1}"] = pad(sin_sg[..., k])

    # --- divergence / del6 factors (geometry.calculate_divg_del6 :500-571, generation._calculate_divg_del6)
    sg = np.stack([out[f"sin_sg{k}"][:, :-1, :-1] for k in range(1, 6)], axis=-1)
    su, sv = out["sina_u"][:, :, :-1], out["sina_v"][:, :-1, :]
    dx_, dy_, dxc_, dyc_ = dx[:, :-1, :], dy[:, :, :-1], dxc[:, :, :-1], dyc[:, :-1, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        divg_u = sv * dyc_ / dx_
        del6_u = sv * dx_ / dyc_
        divg_v = su * dxc_ / dy_
        del6_v = su * dy_ / dxc_
        for jj, ja, jb in ((n, n, n - 1), (-n - 1, -n, -n - 1)):
            f = 0.5 * (sg[:, :, ja, 1] + sg[:, :, jb, 3])
            divg_u[:, :, jj] = f * dyc_[:, :, jj] / dx_[:, :, jj]
            del6_u[:, :, jj] = f * dx_[:, :, jj] / dyc_[:, :, jj]
        for ii, ia, ib in ((n, n, n - 1), (-n - 1, -n, -n - 1)):
            f = 0.5 * (sg[:, ia, :, 0] + sg[:, ib, :, 2])
            divg_v[:, ii, :] = f * dxc_[:, ii, :] / dy_[:, ii, :]
            del6_v[:, ii, :] = f * dy_[:, ii, :] / dxc_[:, ii, :]
    DU, DV, LU, LV = (np.zeros((6, ni, ni)) for _ in range(4))
    DU[:, :-1, :], LU[:, :-1, :] = divg_u, del6_u
    DV[:, :, :-1], LV[:, :, :-1] = divg_v, del6_v
    halo.vector(DV, DU, YD, XD)
    halo.vector(LV, LU, YD, XD)
    for a in (DU, DV, LU, LV):
        np.abs(a, out=a)
    out.update(divg_u=DU, divg_v=DV, del6_u=LU, del6_v=LV)

    # --- lat-lon <-> cubed wind transformation (geometry.unit_vector_lonlat, calculate_grid_z / _a :574-587)
    la, ta = lon_a[:, :-1, :-1], lat_a[:, :-1, :-1]
    vlon = np.stack([-np.sin(la), np.cos(la), np.zeros_like(la)], axis=-1)
    vlat = np.stack([-np.sin(ta) * np.cos(la), -np.sin(ta) * np.sin(la), np.cos(ta)], axis=-1)
    z11, z12 = (ec1 * vlon).sum(-1), (ec1 * vlat).sum(-1)
    z21, z22 = (ec2 * vlon).sum(-1), (ec2 * vlat).sum(-1)
    s5 = sin_sg[..., 4]
    out.update(a11=pad(0.5 * z22 / s5), a12=pad(-0.5 * z12 / s5), a21=pad(-0.5 * z21 / s5), a22=pad(0.5 * z11 / s5))

    # --- a2b edge interpolation factors (geometry.edge_factors :590-700), tile level: corner points 4 .. N+2
    def west_factor(G, Ag, j0, j1):
        py0, py1 = midpoint_ll(Ag[:, n - 1, j0 - 1:j1, 0], Ag[:, n, j0 - 1:j1, 0], Ag[:, n - 1, j0 - 1:j1, 1], Ag[:, n, j0 - 1:j1, 1])
        d1 = gc_dist(py0[:, :-1], G[:, n, j0:j1, 0], py1[:, :-1], G[:, n, j0:j1, 1])
        d2 = gc_dist(py0[:, 1:], G[:, n, j0:j1, 0], py1[:, 1:], G[:, n, j0:j1, 1])
        return d2 / (d1 + d2)

    j0, j1 = 4, N + H
    Ag = agrid[:, :-1, :-1]
    ew_ = np.full((6, ni), BIG)
    ee_ = np.full((6, ni), BIG)
    es_ = np.full((6, ni), BIG)
    en_ = np.full((6, ni), BIG)
    ew_[:, j0:j1] = west_factor(grid, Ag, j0, j1)
    ee_[:, j0:j1] = west_factor(grid[:, ::-1], Ag[:, ::-1], j0, j1)
    es_[:, j0:j1] = west_factor(grid.transpose(0, 2, 1, 3), Ag.transpose(0, 2, 1, 3), j0, j1)
    en_[:, j0:j1] = west_factor(grid[:, :, ::-1].transpose(0, 2, 1, 3), Ag[:, :, ::-1].transpose(0, 2, 1, 3), j0, j1)
    out.update(_edge_w=ew_, _edge_e=ee_, _edge_s=es_, _edge_n=en_)

    # --- Coriolis (helper.py:352-377) and reciprocals (generation.py:1359-1473)
    out["fC"] = 2.0 * c.OMEGA * np.sin(lat)
    out["fC_agrid"] = 2.0 * c.OMEGA * np.sin(lat_a)
    with np.errstate(divide="ignore"):
        for nm in ("dx", "dy", "dxa", "dya", "dxc", "dyc"):
            out["r" + nm] = 1.0 / out[nm]
        out["rarea"] = 1.0 / area
        out["rarea_c"] = 1.0 / area_c
    out["_da_min"] = float(area[:, ce, ce].min())
    out["_da_min_c"] = float(area_c[:, ce, ce].min())
    return out


HORIZONTAL = [
    "dx", "dy", "dxa", "dya", "dxc", "dyc", "rdx", "rdy", "rdxa", "rdya", "rdxc", "rdyc", "area", "area_64", "rarea",
    "rarea_c", "cosa", "cosa_u", "cosa_v", "cosa_s", "sina_u", "sina_v", "rsina", "rsin_u", "rsin_v", "rsin2",
    "sin_sg1", "sin_sg2", "sin_sg3", "sin_sg4", "cos_sg1", "cos_sg2", "cos_sg3", "cos_sg4", "fC", "fC_agrid", "lon",
    "lat", "lon_agrid", "lat_agrid", "a11", "a12", "a21", "a22", "ee1", "ee2", "es1", "ew2",
]


def subdomain_arrays(tiles: Dict[str, np.ndarray], N: int, layout: int, ranks, nz: int = 79) -> List[Dict[str, np.ndarray]]:
    """Cut the tile-level terms into per-subdomain dicts keyed like the reference's GridData dump (per rank)."""
    dec = topology.Decomposition(N // layout, layout)
    n = N // layout
    ak, bk = load_eta(nz)
    p_ref = 1.0e5
    p_int = ak + bk * p_ref
    vert = dict(ak=ak, bk=bk, ptop=np.asarray(ak[0]), p_ref=np.asarray(p_ref),
                dp_ref=ak[1:] - ak[:-1] + (bk[1:] - bk[:-1]) * p_ref,
                p=(p_int[1:] - p_int[:-1]) / np.log(p_int[1:] / p_int[:-1]), ks=np.asarray(int(np.where(bk == 0)[0][-1])))
    res = []
    for r in ranks:
        t = dec.tile_of(r)
        tj, ti = dec.subtile_index(r)
        si = slice(ti * n, ti * n + n + 2 * H + 1)
        sj = slice(tj * n, tj * n + n + 2 * H + 1)
        d = {k: np.ascontiguousarray(tiles[k][t, si, sj]) for k in HORIZONTAL}
        for k in ("divg_u", "divg_v", "del6_u", "del6_v"):
            d["damp_" + k] = np.ascontiguousarray(tiles[k][t, si, sj])
        d["damp_da_min"] = np.asarray(tiles["_da_min"])
        d["damp_da_min_c"] = np.asarray(tiles["_da_min_c"])
        w, e, s, nn = dec.edge_flags(r)
        m = n + 2 * H + 1
        # a2b edge factors exist only on subdomains touching that tile edge; range as geometry.edge_factors
        ew = np.zeros((m, m))
        ee = np.zeros((m, m))
        es = np.zeros(m)
        en = np.zeros(m)
        for a in (ew, ee):
            a[:, H:m - H] = BIG
        for a in (es, en):
            a[H:m - H] = BIG
        gjs, gis = H + tj * n, H + ti * n
        gje = gjs + n - (0 if nn else 1)   # last interface point owned by this subdomain (partitioner.subtile_slice)
        gie = gis + n - (0 if e else 1)
        jst, jen = max(4, gjs) - gjs + H, min(N + H, gje + 2) - gjs + H
        ist, ien = max(4, gis) - gis + H, min(N + H, gie + 2) - gis + H
        if w:
            ew[:, jst:jen] = tiles["_edge_w"][t, sj][None, jst:jen]
        if e:
            ee[:, jst:jen] = tiles["_edge_e"][t, sj][None, jst:jen]
        if s:
            es[ist:ien] = tiles["_edge_s"][t, si][ist:ien]
        if nn:
            en[ist:ien] = tiles["_edge_n"][t, si][ist:ien]
        d.update(edge_w=ew, edge_e=ee, edge_s=es, edge_n=en)
        d.update(vert)
        res.append(d)
    return res


def generate(nx_tile: int, layout: int = 1, ranks=None, nz: int = 79) -> List[Dict[str, np.ndarray]]:
    """Metric terms of the given subdomain ranks (default: all 6*layout^2)."""
    if nx_tile % layout:
        raise ValueError("nx_tile must be divisible by the layout")
    dec = topology.Decomposition(nx_tile // layout, layout)
    ranks = list(range(dec.total_ranks)) if ranks is None else list(ranks)
    return subdomain_arrays(generate_tiles(nx_tile), nx_tile, layout, ranks, nz)
