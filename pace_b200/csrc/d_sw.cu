// D-grid shallow-water Lagrangian dynamics (the hottest stage of the acoustic substep).
//   fv3_d_sw               <- DGridShallowWaterLagrangianDynamics.__call__ (fv3core/pace/fv3core/stencils/d_sw.py:935-1237)
//   fv3_divergence_damping <- DivergenceDamping.__call__ (divergence_damping.py:482-632)
//   fv3_a2b_ord4           <- AGrid2BGridFourthOrder.__call__ (a2b_ord4.py:673-761)
// Sub-stages call the same internal routines as the stand-alone entry points (fxadv.cu, fvtp2d.cu).
#include <type_traits>

#include "a2b.h"
#include "transport.h"

#ifndef FV3_K3B_LOOP
#define FV3_K3B_LOOP true
#endif

namespace fv3 {
int fv_prep_launch(fv3_ctx *ctx, cudaStream_t st, const double *uc, const double *vc, double *crx, double *cry, double *xfx,
                   double *yfx, double *ucc_out, double *vcc_out, double *cx, double *cy, double dt, bool store_all);
}
extern "C" int fv3_fvtp2d(fv3_ctx *, const double *, const double *, const double *, const double *, const double *,
                          double *, double *, const double *, const double *, const double *, int, const double *,
                          const double *, int, int, void *);
extern "C" int fv3_delnflux_nosg(fv3_ctx *, const double *, double *, double *, const double *, const double *, int,
                                 int, void *);

namespace {

constexpr double DCON_THRESHOLD = 1e-5;

struct Ix {
  int isc, iec, jsc, jec, ied, jed;
};
Ix make_ix(const fv3_geom &g) {
  Ix x;
  x.isc = g.halo;
  x.iec = g.halo + g.nx - 1;
  x.jsc = g.halo;
  x.jec = g.halo + g.ny - 1;
  x.ied = x.iec + g.halo;
  x.jed = x.jec + g.halo;
  return x;
}

// fill_corners_bgrid_x/y (corners.py:591-702) as read-time remaps on a B-grid (corner-point) field
FV_HD void bgrid_corner_x(const fv3_geom &g, int s, int &i, int &j) {
  const int isc = g.halo, ic = g.halo + g.nx, jsc = g.halo, jc = g.halo + g.ny;  // ic/jc: east / north corner point
  const bool xo_w = i < isc, xo_e = i > ic, yo_s = j < jsc, yo_n = j > jc;
  if (!((xo_w || xo_e) && (yo_s || yo_n))) return;
  if (!((xo_w ? fv3::on_west(g, s) : fv3::on_east(g, s)) && (yo_s ? fv3::on_south(g, s) : fv3::on_north(g, s)))) return;
  const int a = xo_w ? isc - i : i - ic, b = yo_s ? jsc - j : j - jc;
  i = xo_w ? isc - b : ic + b;
  j = yo_s ? jsc + a : jc - a;
}
FV_HD void bgrid_corner_y(const fv3_geom &g, int s, int &i, int &j) {
  const int isc = g.halo, ic = g.halo + g.nx, jsc = g.halo, jc = g.halo + g.ny;
  const bool xo_w = i < isc, xo_e = i > ic, yo_s = j < jsc, yo_n = j > jc;
  if (!((xo_w || xo_e) && (yo_s || yo_n))) return;
  if (!((xo_w ? fv3::on_west(g, s) : fv3::on_east(g, s)) && (yo_s ? fv3::on_south(g, s) : fv3::on_north(g, s)))) return;
  const int a = xo_w ? isc - i : i - ic, b = yo_s ? jsc - j : j - jc;
  i = xo_w ? isc + b : ic - b;
  j = yo_s ? jsc - a : jc + a;
}

void a2b_ord4_launch(const fv3_ctx *ctx, cudaStream_t st, const double *qin, double *qout, int k0, int k1) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  fv3::launch_planes(ctx, st, k0, k1, 4, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *SQ = b.plane(0), *QX = b.plane(1), *QY = b.plane(2), *OUT = b.plane(3);
    const int64_t ob = O3(s, 0, 0, k);
    fv3::a2b_plane(g, m, s, b, qin + ob, SQ, QX, QY, OUT);
    const int sj2 = g.sj, h2 = g.halo;
    b.rect(h2, h2 + g.nx + 1, b.ja, b.jtop() + 1, [&](int i, int j) {
      const int p = j * sj2 + i;
      qout[ob + p] = OUT[p];
    });
  });
}

// one iteration of the divergence-damping Laplacian (divergence_damping.py:566-589) for levels [k0, nz):
// dnew <- rarea_c * div( grad(dold) scaled by divg_u / divg_v ), cube-corner fills folded in as remaps.
void divg_iteration(const fv3_ctx *ctx, cudaStream_t st, const double *dold, double *dnew, int nt, bool fillc, int k0) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const Ix x = make_ix(g);
  const int isc = x.isc, iec = x.iec, jsc = x.jsc, jec = x.jec;
  fv3::launch3d(ctx, st, isc - nt, iec + nt + 2, jsc - nt, jec + nt + 2, k0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    // away from the cube corners (three points either way) no corner fill is read and no corner term applies:
    // the plain five-point form
    if (!((W || E) && (S || N) && (i <= isc + 2 || i >= iec - 1) && (j <= jsc + 2 || j >= jec - 1))) {
      const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
      const int sj = g.sj;
      const double dc = dold[o];
      const double ucm = (dc - dold[o - sj]) * m.divg_v[o2 - sj], uc0 = (dold[o + sj] - dc) * m.divg_v[o2];
      const double vcm = (dc - dold[o - 1]) * m.divg_u[o2 - 1], vc0 = (dold[o + 1] - dc) * m.divg_u[o2];
      dnew[o] = (ucm - uc0 + vcm - vc0) * m.rarea_c[o2];
      return;
    }
    auto dgx = [&](int ii, int jj) {
      if (fillc) bgrid_corner_x(g, s, ii, jj);
      return dold[O3(s, ii, jj, k)];
    };
    auto dgy = [&](int ii, int jj) {
      if (fillc) bgrid_corner_y(g, s, ii, jj);
      return dold[O3(s, ii, jj, k)];
    };
    auto vc_raw = [&](int ii, int jj) { return (dgx(ii + 1, jj) - dgx(ii, jj)) * m.divg_u[O2(s, ii, jj)]; };
    auto uc_raw = [&](int ii, int jj) { return (dgy(ii, jj + 1) - dgy(ii, jj)) * m.divg_v[O2(s, ii, jj)]; };
    // fill_corners_dgrid_defn(vc, vc, uc, uc, -1) (corners.py:987-1151): x-field = vc (x centre, y interface),
    // y-field = uc (x interface, y centre)
    auto vc_at = [&](int ii, int jj) {
      if (fillc) {
        const bool xo_w = ii < isc, xo_e = ii > iec, yo_s = jj < jsc, yo_n = jj > jec + 1;
        if ((xo_w || xo_e) && (yo_s || yo_n) && (xo_w ? W : E) && (yo_s ? S : N)) {
          const int a = xo_w ? isc - ii : ii - iec, b = yo_s ? jsc - jj : jj - (jec + 1);
          const double sg = (xo_w == yo_s) ? -1.0 : 1.0;  // sw, ne: mysign; nw, se: +1
          const int si = xo_w ? isc - b : iec + 1 + b;
          const int sjj = yo_s ? jsc + a - 1 : jec + 1 - a;
          return sg * uc_raw(si, sjj);
        }
      }
      return vc_raw(ii, jj);
    };
    auto uc_at = [&](int ii, int jj) {
      if (fillc) {
        const bool xo_w = ii < isc, xo_e = ii > iec + 1, yo_s = jj < jsc, yo_n = jj > jec;
        if ((xo_w || xo_e) && (yo_s || yo_n) && (xo_w ? W : E) && (yo_s ? S : N)) {
          const int a = xo_w ? isc - ii : ii - (iec + 1), b = yo_s ? jsc - jj : jj - jec;
          const double sg = (xo_w == yo_s) ? -1.0 : 1.0;
          const int si = xo_w ? isc + b - 1 : iec + 1 - b;
          const int sjj = yo_s ? jsc - a : jec + 1 + a;
          return sg * vc_raw(si, sjj);
        }
      }
      return uc_raw(ii, jj);
    };
    const double ucm = uc_at(i, j - 1), uc0 = uc_at(i, j), vcm = vc_at(i - 1, j), vc0 = vc_at(i, j);
    double d = ucm - uc0 + vcm - vc0;
    const bool ci = (W && i == isc) || (E && i == iec + 1);
    if (ci && S && j == jsc) d = d - ucm;
    if (ci && N && j == jec + 1) d = d + uc0;
    dnew[O3(s, i, j, k)] = d * m.rarea_c[O2(s, i, j)];
  });
}

void divergence_damping(fv3_ctx *ctx, cudaStream_t st, const double *u, const double *v, const double *va, double *vort_b,
                        const double *ua, double *divg_d, const double *vc, const double *uc, double *delpc, double *ke,
                        const double *vort_a, double dt, const fv3_dsw_cols *c) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const Ix x = make_ix(g);
  const int isc = x.isc, iec = x.iec, jsc = x.jsc, jec = x.jec, sj = g.sj, nz = g.nz;
  const int k0 = c->nonzero_nord_k;
  const int nord = c->nonzero_nord;
  const double da_min_c = ctx->c.da_min_c, dddmp = ctx->c.dddmp;
  const double *d2_bg = c->d2_divg;
  if (k0 > 0) {
    // levels [0, k0): second-order damping (divergence_damping.py:21-118,504-548)
    fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, 0, k0, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
      const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
      auto ucd = [&](int ii, int jj) {  // u_contra * dyc at (x centre, y interface)
        const int64_t o = O3(s, ii, jj, k), o2 = O2(s, ii, jj);
        double uco;
        if ((S && jj == jsc) || (N && jj == jec + 1))
          uco = vc[o] > 0 ? u[o] * m.sin_sg4[o2 - sj] : u[o] * m.sin_sg2[o2];
        else
          uco = (u[o] - 0.5 * (va[o - sj] + va[o]) * m.cosa_v[o2]) * m.sina_v[o2];
        return uco * m.dyc[o2];
      };
      auto vcd = [&](int ii, int jj) {
        const int64_t o = O3(s, ii, jj, k), o2 = O2(s, ii, jj);
        double vco;
        if ((W && ii == isc) || (E && ii == iec + 1))
          vco = uc[o] > 0 ? v[o] * m.sin_sg3[o2 - 1] : v[o] * m.sin_sg1[o2];
        else
          vco = (v[o] - 0.5 * (ua[o - 1] + ua[o]) * m.cosa_u[o2]) * m.sina_u[o2];
        return vco * m.dxc[o2];
      };
      const double vm = vcd(i, j - 1), v0 = vcd(i, j), um = ucd(i - 1, j), u0 = ucd(i, j);
      double d = vm - v0 + um - u0;
      const bool ci = (W && i == isc) || (E && i == iec + 1);
      if (ci && S && j == jsc) d = d - vm;
      if (ci && N && j == jec + 1) d = d + v0;
      const int64_t o = O3(s, i, j, k);
      d = m.rarea_c[O2(s, i, j)] * d;
      delpc[o] = d;
      const double delpcdt = d * dt;
      const double damp = da_min_c * fv3::dmax(d2_bg[k], fv3::dmin(0.2, dddmp * fabs(delpcdt)));
      const double vo = damp * d;
      vort_b[o] = vo;
      ke[o] = ke[o] + vo;
    });
  }
  // levels [k0, nz): delpc <- divg_d, then nord Laplacian iterations on divg_d
  fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, k0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    delpc[o] = divg_d[o];
  });
  // the iterations go divg_d -> scratch -> scratch' -> ... and the LAST one writes its (compute-domain) result
  // straight back into divg_d; a single iteration cannot work in place and takes the copy
  double *tmp = fv3::scratch_field(ctx, 7), *tmp2 = fv3::scratch_field(ctx, 8);
  const double *src = divg_d;
  for (int n = 0; n < nord; ++n) {
    const int nt = nord - (n + 1);
    const bool fillc = (n + 1 != nord);
    double *dst = (n + 1 == nord && nord >= 2) ? divg_d : ((n & 1) ? tmp2 : tmp);
    divg_iteration(ctx, st, src, dst, nt, fillc, k0);
    src = dst;
  }
  if (nord == 1) {
    fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, k0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
      const int64_t o = O3(s, i, j, k);
      divg_d[o] = tmp[o];
    });
  }
  const double absdt = fabs(dt);
  double dd8 = 1.0;
  {
    const double base = da_min_c * ctx->c.d4_bg;
    for (int n = 0; n < nord + 1; ++n) dd8 *= base;
    dd8 = pow(base, (double)(nord + 1));
  }
  // a2b_ord4 of the relative vorticity + Smagorinsky-type diffusion + high-order damping
  // (divergence_damping.py:590-632): one plane-resident kernel, the B-grid vorticity never leaves shared memory
  fv3::launch_planes(ctx, st, k0, nz, 4, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *SQ = b.plane(0), *QX = b.plane(1), *QY = b.plane(2), *OUT = b.plane(3);
    const int64_t ob = O3(s, 0, 0, k);
    const bool smag = !(dddmp < 1e-5);
    if (smag) b.prefetch_next_wave(vort_a, g, k);
    if (smag) fv3::a2b_plane(g, m, s, b, vort_a + ob, SQ, QX, QY, OUT);
    const int sj2 = g.sj, h2 = g.halo;
    const double *rarea_unused = nullptr;
    (void)rarea_unused;
    b.rect(h2, h2 + g.nx + 1, b.ja, b.jtop() + 1, [&](int i, int j) {
      const int p = j * sj2 + i;
      const int64_t o = ob + p;
      double vo;
      if (!smag) {
        vo = 0.0;
      } else {
        const double vb = OUT[p];
        const double dp = delpc[o];
        vo = absdt * sqrt(dp * dp + vb * vb);
      }
      const double damp = da_min_c * fv3::dmax(d2_bg[k], fv3::dmin(0.2, dddmp * fabs(vo)));
      vo = damp * delpc[o] + dd8 * divg_d[o];
      vort_b[o] = vo;
      ke[o] = ke[o] + vo;
    });
  });
}

// D-grid wind advected along its own direction to a cell corner: advect_u_along_x / advect_v_along_y
// (xtp_u.py:9-91, ytp_v.py), iord < 8.  q(n): wind at index n along the sweep; dxe(n): dx (or dy) for the edge
// formula of compute_al; zero(n): bl = br = 0 at the cells next to a cube corner (xtp_u.py:39-47)
template <class Q, class DXE, class Z>
FV_HD double advect_along(int mord, Q q, DXE dxe, Z zero, double ub, double cfl, int i, const fv3::Edge1D &e) {
  const double al0 = fv3::ppm_al_lt8(q, dxe, i - 1, e), al1 = fv3::ppm_al_lt8(q, dxe, i, e), al2 = fv3::ppm_al_lt8(q, dxe, i + 1, e);
  const double ql = q(i - 1), qr = q(i);
  double bl_l = al0 - ql, br_l = al1 - ql, bl_r = al1 - qr, br_r = al2 - qr;
  if (zero(i - 1)) bl_l = br_l = 0.0;
  if (zero(i)) bl_r = br_r = 0.0;
  const double b0_l = bl_l + br_l, b0_r = bl_r + br_r;
  const double fx0 = fv3::ppm_fx1(cfl, br_l, b0_l, bl_r, b0_r);
  bool s_l, s_r;
  if (mord == 5) {
    s_l = bl_l * br_l < 0;
    s_r = bl_r * br_r < 0;
  } else {
    s_l = (3.0 * fabs(b0_l)) < fabs(bl_l - br_l);
    s_r = (3.0 * fabs(b0_r)) < fabs(bl_r - br_r);
  }
  const double mask = (s_l || s_r) ? 1.0 : 0.0;
  return ub > 0.0 ? ql + fx0 * mask : qr + fx0 * mask;
}


// kinetic energy at the cell corner (i, j) (d_sw.py:204-298): winds advected along their own direction to the corner
// (xtp_u.py / ytp_v.py), edge and corner forms next to the cube-tile edges
FV_DEV double ke_point(const fv3_geom &g, const fv3_grid &m, int s, int i, int j, int k, const double *u, const double *v,
                      const double *uc, const double *vc, const double *ucc, const double *vcc, double dt, int mord) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1, sj = g.sj;
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    const bool ie_ = (W && i == isc) || (E && i == iec + 1), je_ = (S && j == jsc) || (N && j == jec + 1);
    const double ub_cov = 0.5 * (uc[o - sj] + uc[o]), vb_cov = 0.5 * (vc[o - 1] + vc[o]);
    double ub = (ub_cov - vb_cov * m.cosa[o2]) * m.rsina[o2];
    double vb = (vb_cov - ub_cov * m.cosa[o2]) * m.rsina[o2];
    if (je_) ub = 0.25 * (-ucc[o - 2 * sj] + 3.0 * (ucc[o - sj] + ucc[o]) - ucc[o + sj]);
    if (ie_) ub = 0.5 * (ucc[o - sj] + ucc[o]);
    if (ie_) vb = 0.25 * (-vcc[o - 2] + 3.0 * (vcc[o - 1] + vcc[o]) - vcc[o + 1]);
    if (je_) vb = 0.5 * (vcc[o - 1] + vcc[o]);
    double kev;
    if (ie_ && je_) {
      // corner_ke (d_sw.py:258-283)
      int io1, jo1, io2;
      double vsign;
      if (i == isc && j == jsc) { io1 = 0; jo1 = 0; io2 = -1; vsign = 1; }
      else if (i != isc && j == jsc) { io1 = -1; jo1 = 0; io2 = 0; vsign = -1; }
      else if (i != isc && j != jsc) { io1 = -1; jo1 = -1; io2 = 0; vsign = 1; }
      else { io1 = 0; jo1 = -1; io2 = -1; vsign = -1; }
      const double dt6 = dt / 6.0;
      const double u0 = u[o], um = u[o - 1], v0 = v[o], vm = v[o - sj];
      const double ut0 = ucc[o], utm = ucc[o - sj], vt0 = vcc[o], vtm = vcc[o - 1];
      kev = dt6 * ((ut0 + utm) * ((io1 + 1) * u0 - (io1 * um)) + (vt0 + vtm) * ((jo1 + 1) * v0 - (jo1 * vm)) +
                   (((jo1 + 1) * ut0 - (jo1 * utm)) + vsign * ((io1 + 1) * vt0 - (io1 * vtm))) * ((io2 + 1) * u0 - (io2 * um)));
    } else if (!((W && i <= isc + 3) || (E && i >= iec - 2) || (S && j <= jsc + 3) || (N && j >= jec - 2))) {
      // away from the tile edges: interior edge values, no zeroed parabolas (same expressions as advect_along)
      auto adv = [&](const double *q, int64_t st, double ubv, double cfl) {
        const double qm2 = q[-2 * st], ql = q[-st], qr = q[0], qp1 = q[st];
        const double al0 = fv3::PPM_P1 * (qm2 + ql) + fv3::PPM_P2 * (q[-3 * st] + qr);
        const double al1 = fv3::PPM_P1 * (ql + qr) + fv3::PPM_P2 * (qm2 + qp1);
        const double al2 = fv3::PPM_P1 * (qr + qp1) + fv3::PPM_P2 * (ql + q[2 * st]);
        const double bl_l = al0 - ql, br_l = al1 - ql, bl_r = al1 - qr, br_r = al2 - qr;
        const double b0_l = bl_l + br_l, b0_r = bl_r + br_r;
        const double fx0 = fv3::ppm_fx1(cfl, br_l, b0_l, bl_r, b0_r);
        bool s_l, s_r;
        if (mord == 5) {
          s_l = bl_l * br_l < 0;
          s_r = bl_r * br_r < 0;
        } else {
          s_l = (3.0 * fabs(b0_l)) < fabs(bl_l - br_l);
          s_r = (3.0 * fabs(b0_r)) < fabs(bl_r - br_r);
        }
        const double mask = (s_l || s_r) ? 1.0 : 0.0;
        return ubv > 0.0 ? ql + fx0 * mask : qr + fx0 * mask;
      };
      const double cflx = ub > 0 ? ub * dt * m.rdx[o2 - 1] : ub * dt * m.rdx[o2];
      const double cfly = vb > 0 ? vb * dt * m.rdy[o2 - sj] : vb * dt * m.rdy[o2];
      kev = 0.5 * dt * (ub * adv(u + o, 1, ub, cflx) + vb * adv(v + o, sj, vb, cfly));
    } else {
      const fv3::Edge1D ex{W, E, isc, iec}, ey{S, N, jsc, jec};
      auto qu = [&](int ii) { return u[O3(s, ii, j, k)]; };
      auto dxe = [&](int ii) { return m.dx[O2(s, ii, j)]; };
      auto zx = [&](int ii) { return je_ && ((W && (ii == isc - 1 || ii == isc)) || (E && (ii == iec || ii == iec + 1))); };
      const double cflx = ub > 0 ? ub * dt * m.rdx[o2 - 1] : ub * dt * m.rdx[o2];
      const double adv_u = advect_along(mord, qu, dxe, zx, ub, cflx, i, ex);
      auto qv = [&](int jj) { return v[O3(s, i, jj, k)]; };
      auto dye = [&](int jj) { return m.dy[O2(s, i, jj)]; };
      auto zy = [&](int jj) { return ie_ && ((S && (jj == jsc - 1 || jj == jsc)) || (N && (jj == jec || jj == jec + 1))); };
      const double cfly = vb > 0 ? vb * dt * m.rdy[o2 - sj] : vb * dt * m.rdy[o2];
      const double adv_v = advect_along(mord, qv, dye, zy, vb, cfly, j, ey);
      kev = 0.5 * dt * (ub * adv_u + vb * adv_v);
    }
    return kev;
}

// ---- plane forms of the divergence-damping pieces, used by the fused momentum stage -----------------------------------
// second-order damping divergence of the sponge levels at corner (i, j) (divergence_damping.py:21-118)
FV_DEV double div2_point(const fv3_geom &g, const fv3_grid &m, int s, int i, int j, int k, const double *u, const double *v,
                        const double *ua, const double *va, const double *uc, const double *vc) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1, sj = g.sj;
  const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
  auto ucd = [&](int ii, int jj) {  // u_contra * dyc at (x centre, y interface)
    const int64_t o = O3(s, ii, jj, k), o2 = O2(s, ii, jj);
    double uco;
    if ((S && jj == jsc) || (N && jj == jec + 1))
      uco = vc[o] > 0 ? u[o] * m.sin_sg4[o2 - sj] : u[o] * m.sin_sg2[o2];
    else
      uco = (u[o] - 0.5 * (va[o - sj] + va[o]) * m.cosa_v[o2]) * m.sina_v[o2];
    return uco * m.dyc[o2];
  };
  auto vcd = [&](int ii, int jj) {
    const int64_t o = O3(s, ii, jj, k), o2 = O2(s, ii, jj);
    double vco;
    if ((W && ii == isc) || (E && ii == iec + 1))
      vco = uc[o] > 0 ? v[o] * m.sin_sg3[o2 - 1] : v[o] * m.sin_sg1[o2];
    else
      vco = (v[o] - 0.5 * (ua[o - 1] + ua[o]) * m.cosa_u[o2]) * m.sina_u[o2];
    return vco * m.dxc[o2];
  };
  const double vm = vcd(i, j - 1), v0 = vcd(i, j), um = ucd(i - 1, j), u0 = ucd(i, j);
  double d = vm - v0 + um - u0;
  const bool ci = (W && i == isc) || (E && i == iec + 1);
  if (ci && S && j == jsc) d = d - vm;
  if (ci && N && j == jec + 1) d = d + v0;
  return m.rarea_c[O2(s, i, j)] * d;
}

// one Laplacian iteration of the divergence damping (divergence_damping.py:566-589) on shared planes: DST <- rarea_c *
// div(grad(SRC) scaled by divg_u / divg_v) on corners [isc-nt, iec+nt+1] x [j0, j1), cube-corner fills as read remaps
FV_DEV void divg_iter_plane(const fv3_geom &g, const fv3_grid &m, int s, const fv3::Block &b, const double *SRC, double *DST,
                           int nt, bool fillc, int j0, int j1) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1, sj = g.sj;
  const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
  const int64_t o2b = O2(s, 0, 0);
  const double *divg_u = m.divg_u + o2b, *divg_v = m.divg_v + o2b, *rarea_c = m.rarea_c + o2b;
  b.rect(isc - nt, iec + nt + 2, j0, j1, [&](int i, int j) {
    const int p = j * sj + i;
    if (!((W || E) && (S || N) && (i <= isc + 2 || i >= iec - 1) && (j <= jsc + 2 || j >= jec - 1))) {
      const double dc = SRC[p];
      const double ucm = (dc - SRC[p - sj]) * divg_v[p - sj], uc0 = (SRC[p + sj] - dc) * divg_v[p];
      const double vcm = (dc - SRC[p - 1]) * divg_u[p - 1], vc0 = (SRC[p + 1] - dc) * divg_u[p];
      DST[p] = (ucm - uc0 + vcm - vc0) * rarea_c[p];
      return;
    }
    auto dgx = [&](int ii, int jj) {
      if (fillc) bgrid_corner_x(g, s, ii, jj);
      return SRC[jj * sj + ii];
    };
    auto dgy = [&](int ii, int jj) {
      if (fillc) bgrid_corner_y(g, s, ii, jj);
      return SRC[jj * sj + ii];
    };
    auto vc_raw = [&](int ii, int jj) { return (dgx(ii + 1, jj) - dgx(ii, jj)) * divg_u[jj * sj + ii]; };
    auto uc_raw = [&](int ii, int jj) { return (dgy(ii, jj + 1) - dgy(ii, jj)) * divg_v[jj * sj + ii]; };
    auto vc_at = [&](int ii, int jj) {
      if (fillc) {
        const bool xo_w = ii < isc, xo_e = ii > iec, yo_s = jj < jsc, yo_n = jj > jec + 1;
        if ((xo_w || xo_e) && (yo_s || yo_n) && (xo_w ? W : E) && (yo_s ? S : N)) {
          const int a = xo_w ? isc - ii : ii - iec, bb = yo_s ? jsc - jj : jj - (jec + 1);
          const double sg = (xo_w == yo_s) ? -1.0 : 1.0;
          const int si = xo_w ? isc - bb : iec + 1 + bb;
          const int sjj = yo_s ? jsc + a - 1 : jec + 1 - a;
          return sg * uc_raw(si, sjj);
        }
      }
      return vc_raw(ii, jj);
    };
    auto uc_at = [&](int ii, int jj) {
      if (fillc) {
        const bool xo_w = ii < isc, xo_e = ii > iec + 1, yo_s = jj < jsc, yo_n = jj > jec;
        if ((xo_w || xo_e) && (yo_s || yo_n) && (xo_w ? W : E) && (yo_s ? S : N)) {
          const int a = xo_w ? isc - ii : ii - (iec + 1), bb = yo_s ? jsc - jj : jj - jec;
          const double sg = (xo_w == yo_s) ? -1.0 : 1.0;
          const int si = xo_w ? isc + bb - 1 : iec + 1 - bb;
          const int sjj = yo_s ? jsc - a : jec + 1 + a;
          return sg * vc_raw(si, sjj);
        }
      }
      return uc_raw(ii, jj);
    };
    const double ucm = uc_at(i, j - 1), uc0 = uc_at(i, j), vcm = vc_at(i - 1, j), vc0 = vc_at(i, j);
    double d = ucm - uc0 + vcm - vc0;
    const bool ci = (W && i == isc) || (E && i == iec + 1);
    if (ci && S && j == jsc) d = d - ucm;
    if (ci && N && j == jec + 1) d = d + uc0;
    DST[p] = d * rarea_c[p];
  });
}

// ---- K3a: kinetic energy, relative vorticity and the divergence damping of d_sw (d_sw.py:204-328,
// divergence_damping.py:482-632), ONE strip-resident kernel.  The relative vorticity plane, the Laplacian iterates of
// the divergence and the three A->B interpolation temporaries live in shared memory; written: ke (+ damping), the
// damped B-grid vorticity, and the A-grid vorticity (read again by K3b's transport and del-n damping).  delpc and
// divg_d are scratch in the reference after this point (nothing downstream reads them) and are not written.
int dsw_vorticity_launch(fv3_ctx *ctx, cudaStream_t st, const double *u, const double *v, const double *ua, const double *va,
                         const double *uc, const double *vc, const double *ucc, const double *vcc, const double *divgd,
                         double *ke, double *vort_a, double *vort_b, double dt, const fv3_dsw_cols *c) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const fv3_config cfg = ctx->c;
  const int k0 = c->nonzero_nord_k, nord = c->nonzero_nord;
  const double da_min_c = cfg.da_min_c, dddmp = cfg.dddmp;
  const double *d2_bg = c->d2_divg;
  const int mord = cfg.hord_mt < 0 ? -cfg.hord_mt : cfg.hord_mt;
  const double absdt = fabs(dt);
  const double dd8 = pow(da_min_c * cfg.d4_bg, (double)(nord + 1));
  return fv3::launch_planes(ctx, st, 0, g.nz, 5, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *VA = b.plane(0), *DF = b.plane(1), *QX = b.plane(2), *QY = b.plane(3), *OUT = b.plane(4);
    const int sj = g.sj, h = g.halo;
    const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1, ied = iec + h, jed = jec + h;
    const int ja = b.ja, jt = b.jtop();
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const bool sponge = k < k0;
    if (!sponge && nord > 0) {  // the divergence of the previous C-grid step: start its transfer now
      b.bulk_begin(1, sj);
      b.bulk_rows(QX, divgd + ob, sj);
    }
    // relative vorticity on the A grid (d_sw.py:301-328) on every resident row; stored for the rows this strip owns
    {
      const double *dxm = m.dx + o2b, *dym = m.dy + o2b, *rarea = m.rarea + o2b;
      const double *up = u + ob, *vp = v + ob;
      const int jlo = b.first ? 0 : ja, jhi = b.last ? jed + 1 : b.jb;
      b.rect(0, ied + 1, b.lo(0, h), b.hi(jed + 1, h), [&](int i, int j) {
        const int p = j * sj + i;
        const double rdy_tmp = rarea[p] * dxm[p], rdx_tmp = rarea[p] * dym[p];
        const double vo = (up[p] - up[p + sj] * dxm[p + sj] / dxm[p]) * rdy_tmp + (vp[p + 1] * dym[p + 1] / dym[p] - vp[p]) * rdx_tmp;
        VA[p] = vo;
        if (j >= jlo && j < jhi) vort_a[ob + p] = vo;
      });
    }
    const bool smag = !(dddmp < 1e-5);
    if (!sponge) {
      // nord Laplacian iterations of the divergence (divergence_damping.py:566-589), ping-pong between the planes the
      // A->B interpolation uses afterwards; the last iterate lands in DF
      if (nord > 0) b.bulk_wait();
      const double *src = QX;
      for (int n = 0; n < nord; ++n) {
        const int nt = nord - (n + 1);
        double *dst = (n + 1 == nord) ? DF : (src == QX ? QY : QX);
        divg_iter_plane(g, m, s, b, src, dst, nt, n + 1 != nord, b.lo(jsc - nt, nt), b.hi(jec + nt + 2, nt + 1));
        src = dst;
      }
      if (smag) fv3::a2b_plane(g, m, s, b, vort_a + ob, VA, QX, QY, OUT, true);
    }
    // damping term and kinetic energy on the corners this strip owns
    b.rect(isc, iec + 2, ja, jt + 1, [&](int i, int j) {
      const int p = j * sj + i;
      const int64_t o = ob + p;
      double vo;
      if (sponge) {
        const double d = div2_point(g, m, s, i, j, k, u, v, ua, va, uc, vc);
        const double damp = da_min_c * fv3::dmax(d2_bg[k], fv3::dmin(0.2, dddmp * fabs(d * dt)));
        vo = damp * d;
      } else {
        const double dp = divgd[o];
        double vo0 = 0.0;
        if (smag) {
          const double vb = OUT[p];
          vo0 = absdt * sqrt(dp * dp + vb * vb);
        }
        const double damp = da_min_c * fv3::dmax(d2_bg[k], fv3::dmin(0.2, dddmp * fabs(vo0)));
        vo = damp * dp + dd8 * (nord > 0 ? DF[p] : dp);
      }
      vort_b[o] = vo;
      ke[o] = ke_point(g, m, s, i, j, k, u, v, uc, vc, ucc, vcc, dt, mord) + vo;
    });
  });
}

// ---- K3b: absolute-vorticity transport, wind update from the kinetic-energy gradient and the vorticity flux, del-n
// damping of the vorticity, the heat it dissipates, final winds (d_sw.py:1131-1237), ONE strip-resident kernel.
// u's first row of every strip but the first is parked (the strip below still reads the old value), see unpark below.
template <int MORD>
int dsw_winds_launch(fv3_ctx *ctx, cudaStream_t st, double *u, double *v, const double *ke, const double *vort_a,
                     const double *vort_b, const double *crx, const double *cry, const double *xfx, const double *yfx,
                     const double *delp, const double *heat_s, double *heat_source, double *side_u, double dt,
                     const fv3_dsw_cols *c) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const fv3_config cfg = ctx->c;
  const fv3_dsw_cols cl = *c;
  const double d_con_cfg = cfg.d_con;
  return fv3::launch_planes(ctx, st, 0, g.nz, fv3::FVTP_PLANES, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.plane(0), *A = b.plane(1), *B = b.plane(2), *D = b.plane(3), *T = b.plane(4);
    const int sj = g.sj, h = g.halo, nx = g.nx;
    const int isc = h, iec = h + nx - 1;
    const int ja = b.ja, jb = b.jb, jt = b.jtop();
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const double *dxm = m.dx + o2b, *dym = m.dy + o2b, *rdx = m.rdx + o2b, *rdy = m.rdy + o2b;
    // absolute vorticity = relative vorticity + f0, transported with the area fluxes (d_sw.py:1131-1147)
    fv3::PlaneArgs pa{vort_a, crx, cry, xfx, yfx, xfx, yfx};
    pa.add2d = m.f0 + o2b;
    fv3::fvtp2d_plane<MORD, true, false, FV3_K3B_LOOP>(g, m, s, k, b, pa, Q, A, B, D, T);
    // u_and_v_from_ke (d_sw.py:439-477): u' -> Q on faces [ja, jb], v' -> D on rows [ja, jb)
    {
      const double *kp = ke + ob, *up = u + ob, *vp = v + ob;
      b.rect(isc, iec + 2, ja, jb + 1, [&](int i, int j) {
        const int p = j * sj + i;
        if (i <= iec) Q[p] = up[p] * dxm[p] + kp[p] - kp[p + 1] + A[p];
        if (j < jb) D[p] = vp[p] * dym[p] + kp[p] - kp[p + sj] - B[p];
      });
    }
    // del-n damping fluxes of the relative vorticity (d_sw.py:1160-1166): fx2 -> A ("ut"), fy2 -> B ("vt")
    fv3::delnflux_plane(g, m, s, b, vort_a + ob, cl.dn_damp_vt_c[k], cl.nord_v[k] > 0, cl.nmax_v, false, T, A, B);
    // vort_differencing + heat_source_from_vorticity_damping (d_sw.py:349-577) on the cells this strip owns
    {
      const bool dc = cl.d_con[k] > DCON_THRESHOLD;
      const double dcon_k = cl.d_con[k];
      const double *vb_ = vort_b + ob, *rsin2 = m.rsin2 + o2b, *cosa_s = m.cosa_s + o2b;
      b.rect(isc, iec + 1, ja, jb, [&](int i, int j) {
        const int p = j * sj + i;
        const int64_t o = ob + p;
        double hs = heat_s[o];
        if (dc) {
          auto ubt = [&](int pp) { return ((vb_[pp] - vb_[pp + 1]) + B[pp]) * rdx[pp]; };
          auto vbt = [&](int pp) { return ((vb_[pp] - vb_[pp + sj]) - A[pp]) * rdy[pp]; };
          const double ub0 = ubt(p), ub1 = ubt(p + sj), vb0 = vbt(p), vb1 = vbt(p + 1);
          const double fy0 = Q[p] * rdx[p], fy1 = Q[p + sj] * rdx[p + sj];
          const double fx0 = D[p] * rdy[p], fx1 = D[p + 1] * rdy[p + 1];
          const double gy0 = fy0 * ub0, gy1 = fy1 * ub1, gx0 = fx0 * vb0, gx1 = fx1 * vb1;
          const double u2 = fy0 + fy1, du2 = ub0 + ub1, v2 = fx0 + fx1, dv2 = vb0 + vb1;
          const double dampterm = rsin2[p] * 0.25 *
                                  ((ub0 * ub0 + ub1 * ub1 + vb0 * vb0 + vb1 * vb1) + 2.0 * (gy0 + gy1 + gx0 + gx1) -
                                   cosa_s[p] * (u2 * dv2 + v2 * du2 + du2 * dv2));
          hs = delp[o] * (hs - dcon_k * dampterm);
        }
        if (d_con_cfg > DCON_THRESHOLD) heat_source[o] = heat_source[o] + hs;
      });
    }
    // update_u_and_v (d_sw.py:582-608) and the stores
    {
      const bool dv = cl.damp_vt[k] > 1e-5;
      b.rect(isc, iec + 2, ja, jt + 1, [&](int i, int j) {
        const int p = j * sj + i;
        if (i <= iec) {
          const double un = dv ? Q[p] + B[p] : Q[p];
          ((j == ja && !b.first) ? side_u : u)[ob + p] = un;
        }
        if (j < jb) v[ob + p] = dv ? D[p] - A[p] : D[p];
      });
    }
  });
}

struct Fields4 {
  double *f[4];
};

// Rows of an in-place update that another strip of the same plane still reads as halo are written to a side buffer
// (plane-for-plane copy of the field) and copied back by this launch once every strip is done.
void unpark_rows(const fv3_ctx *ctx, cudaStream_t st, Fields4 fl, int nf, double *side0, int nk) {
  const fv3_geom g = ctx->g;
  const fv3::StripGeom sg = fv3::strip_geometry(g, fv3::FVTP_PLANES, nk);
  if (sg.ns <= 1) return;
  const int64_t side_stride = g.ss * g.n_sub;
  const int h = g.halo, R = sg.rows_per_strip;
  fv3::launch3d(ctx, st, h, h + g.nx, 0, 2 * h * (sg.ns - 1), 0, nk, FV_LAMBDA(int s, int i, int jj, int k) { FV_DEV_GM
    const int h2 = g.halo, bnd = jj / (2 * h2), j = h2 + (bnd + 1) * R - h2 + (jj - bnd * 2 * h2);
    if (j >= h2 + g.ny) return;
    const int64_t o = O3(s, i, j, k);
    for (int n = 0; n < nf; ++n) fl.f[n][o] = side0[n * side_stride + o];
  });
}

// ---- K2: the four flux-form transports of d_sw (delp, w, q_con, pt) and everything between them, strip-resident
// (d_sw.py:967-1090): delp transport + del-n damping -> mass fluxes (kept in a scratch field that stays in L2,
// accumulated into mfx / mfy: flux_capacitor :29-50); del-n fluxes of damp_w * w and the heat they dissipate (:53-103);
// w, q_con, pt transports with the mass fluxes, their del-n damping, and the flux-form updates of all four fields
// (apply_fluxes, apply_pt_delp_fluxes, adjust_w_and_qcon :106-160,331-346).
// ONE source, instantiated per PARTS (bit mask: 1 delp, 2 w, 4 q_con, 8 pt and the new mass): the stage runs as four
// kernels of 7.7-8.5 K SASS instructions (FV3_K2_SPLIT = 4) — measured faster than two (PARTS 3, 12) or one (15, 38 K
// instructions) although every kernel stages crx / cry / xfx / yfx / delp again; the mass fluxes and the old delp are
// read-only after the first part, so the parts only have to run in order.  Within a part the sequence is straight-line
// (field pointers and column parameters stay kernel-parameter constants).
// MDP / MVT / MTM: |hord_dp|, |hord_vt|, |hord_tm| (5, 6 or 8).
template <int MDP, int MVT, int MTM, int PARTS = 15>
int dsw_scalars_launch(fv3_ctx *ctx, cudaStream_t st, double *delp, double *pt, double *w, double *q_con, const double *crx,
                       const double *cry, const double *xfx, const double *yfx, double *mfx, double *mfy, double *heat_s,
                       double *diss_est, double *fxs, double *fys, double *side0, double dt, const fv3_dsw_cols *c,
                       int rdp, int rvt, int rtm) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const fv3_dsw_cols cl = *c;
  const int64_t side_stride = g.ss * g.n_sub;
  return fv3::launch_planes(ctx, st, 0, g.nz, fv3::FVTP_PLANES, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.plane(0), *A = b.plane(1), *B = b.plane(2), *D = b.plane(3), *T = b.plane(4);
    const int sj = g.sj, h = g.halo, nx = g.nx;
    const int isc = h, iec = h + nx - 1;
    const int ja = b.ja, jb = b.jb, jt = b.jtop();
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const double *rarea = m.rarea + o2b;
    const double *dp = delp + ob;
    double *fxp = fxs + ob, *fyp = fys + ob;
    const double damp_w = cl.damp_w[k];
    // rows another strip of this plane reads as halo are parked (see unpark_rows)
    auto parked = [&](int j) { return (j < ja + h && !b.first) || (j >= jb - h && !b.last); };
    // new mass of a cell (apply_pt_delp_fluxes)
    auto dp_new = [&](int p, double dp0) { return dp0 + (fxp[p] - fxp[p + 1] + fyp[p] - fyp[p + sj]) * rarea[p]; };
    // Straight-line sequence (field pointers and column parameters stay kernel-parameter constants instead of
    // loop-carried registers): delp | w | q_con | pt.
    auto transport = [&](auto mord_tag, const double *q, const double *xu, const double *yu) {
      constexpr int MO = decltype(mord_tag)::value;
      const fv3::PlaneArgs pa{q, crx, cry, xfx, yfx, xu, yu};
      if (MO != 0) {
        fv3::fvtp2d_plane<MO == 0 ? 6 : MO, false, false, true>(g, m, s, k, b, pa, Q, A, B, D, T);
      } else {
        // mixed orders: the true |hord| of this field is a run-time value
        const int mord = q == w ? rvt : (q == pt ? rtm : rdp);
        if (mord == 8)
          fv3::fvtp2d_plane<8, false, false, true>(g, m, s, k, b, pa, Q, A, B, D, T);
        else if (mord == 5)
          fv3::fvtp2d_plane<5, false, false, true>(g, m, s, k, b, pa, Q, A, B, D, T);
        else
          fv3::fvtp2d_plane<6, false, false, true>(g, m, s, k, b, pa, Q, A, B, D, T);
      }
    };
    constexpr bool same = MDP == MVT && MDP == MTM;
    typedef std::integral_constant<int, same ? MDP : 0> TagDP;
    typedef std::integral_constant<int, same ? MVT : 0> TagVT;
    typedef std::integral_constant<int, same ? MTM : 0> TagTM;
    // (1) mass: delp transport + del-n damping (d_sw.py:967-975); mass fluxes kept for the other transports and
    //     accumulated (flux_capacitor)
    if (PARTS & 1) {
    transport(TagDP(), delp, xfx, yfx);
    fv3::delnflux_plane(g, m, s, b, dp, cl.dn_damp_vt[k], cl.nord_v[k] > 0, cl.nmax_v, false, Q, D, T);
    b.rect(isc, iec + 2, ja, jb + 1, [&](int i, int j) {
      const int p = j * sj + i;
      if (j < jb) {
        const double fl = B[p] + D[p];
        fxp[p] = fl;
        mfx[ob + p] = mfx[ob + p] + fl;
      }
      if (i <= iec) {
        const double fl = A[p] + T[p];
        fyp[p] = fl;  // face row jb is also the next strip's first face: both store the same value
        if (j <= jt) mfy[ob + p] = mfy[ob + p] + fl;
      }
    });
    }
    if (PARTS & 2) {
    // (2) w: del-n fluxes of damp_w * w and the heat they dissipate (d_sw.py:53-103); dw goes to T for the update
    fv3::delnflux_plane(g, m, s, b, w + ob, cl.dn_damp_w_c[k], cl.nord_w[k] > 0, cl.nmax_w, false, Q, A, B);
    {
      const double dd8 = cl.ke_bg[k] * fabs(dt);
      b.rect(isc, iec + 1, ja, jb, [&](int i, int j) {
        const int p = j * sj + i;
        double hs = 0.0;
        if (damp_w > 1e-5) {
          const double d = (A[p] - A[p + 1] + B[p] - B[p + sj]) * rarea[p];
          T[p] = d;
          hs = dd8 - d * (w[ob + p] + 0.5 * d);
        }
        heat_s[ob + p] = hs;
        diss_est[ob + p] = hs;
      });
    }
    // (3) w transport and update (apply_fluxes + adjust_w_and_qcon)
    transport(TagVT(), w, fxs, fys);
    b.rect(isc, iec + 1, ja, jb, [&](int i, int j) {
      const int p = j * sj + i;
      const int64_t o = ob + p;
      const double dp0 = dp[p], ra = rarea[p];
      double wv = w[o] * dp0 + (B[p] - B[p + 1] + A[p] - A[p + sj]) * ra;
      wv = wv / dp_new(p, dp0);
      if (damp_w > 1e-5) wv = wv + T[p];
      (parked(j) ? side0 : w)[o] = wv;
    });
    }
    // (4) q_con and (5) pt: transport, mass-weighted del-n damping (delnflux.py:1164-1207, 215-238), update
    auto damped_update = [&](double *q, double dk, bool hi, int nmax, double *side) {
      fv3::delnflux_plane(g, m, s, b, q + ob, dk, hi, nmax, true, Q, D, T);
      b.rect(isc, iec + 2, ja, jb + 1, [&](int i, int j) {
        const int p = j * sj + i;
        if (j < jb) B[p] = B[p] + 0.5 * dk * (dp[p - 1] + dp[p]) * D[p];
        if (i <= iec) A[p] = A[p] + 0.5 * dk * (dp[p - sj] + dp[p]) * T[p];
      });
      b.rect(isc, iec + 1, ja, jb, [&](int i, int j) {
        const int p = j * sj + i;
        const int64_t o = ob + p;
        const double dp0 = dp[p];
        double qv = q[o] * dp0 + (B[p] - B[p + 1] + A[p] - A[p + sj]) * rarea[p];
        qv = qv / dp_new(p, dp0);
        (parked(j) ? side : q)[o] = qv;
      });
    };
    if (PARTS & 4) {
      transport(TagDP(), q_con, fxs, fys);
      damped_update(q_con, cl.dn_damp_t[k], cl.nord_t[k] > 0, cl.nmax_t, side0 + side_stride);
    }
    if (PARTS & 8) {
      transport(TagTM(), pt, fxs, fys);
      damped_update(pt, cl.dn_damp_vt[k], cl.nord_v[k] > 0, cl.nmax_v, side0 + 2 * side_stride);
    }
    // the new mass itself
    if (PARTS & 8) {
      double *side = side0 + 3 * side_stride;
      b.rect(isc, iec + 1, ja, jb, [&](int i, int j) {
        const int p = j * sj + i;
        (parked(j) ? side : delp)[ob + p] = dp_new(p, dp[p]);
      });
    }
  });
}

}  // namespace

extern "C" {

int fv3_a2b_ord4(fv3_ctx *ctx, const double *qin, double *qout, int kstart, int nk, void *stream) {
  a2b_ord4_launch(ctx, (cudaStream_t)stream, qin, qout, kstart, kstart + nk);
  return fv3::check_launch("fv3_a2b_ord4");
}

int fv3_divergence_damping(fv3_ctx *ctx, const double *u, const double *v, const double *va, double *damped_rel_vort_bgrid,
                           const double *ua, double *divg_d, const double *vc, const double *uc, double *delpc, double *ke,
                           const double *rel_vort_agrid, double dt, const fv3_dsw_cols *cols, void *stream) {
  if (cols->nonzero_nord > 3) {
    fv3::set_error("fv3_divergence_damping: nord > 3");
    return -1;
  }
  divergence_damping(ctx, (cudaStream_t)stream, u, v, va, damped_rel_vort_bgrid, ua, divg_d, vc, uc, delpc, ke,
                     rel_vort_agrid, dt, cols);
  return fv3::check_launch("fv3_divergence_damping");
}

int fv3_d_sw(fv3_ctx *ctx, double *delpc, double *delp, double *pt, double *u, double *v, double *w, double *uc,
             double *vc, const double *ua, const double *va, double *divgd, double *mfx, double *mfy, double *cx,
             double *cy, double *crx, double *cry, double *xfx, double *yfx, double *q_con, const double *zh,
             double *heat_source, double *diss_est, double dt, const fv3_dsw_cols *c, void *stream) {
  (void)zh;
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const fv3_config cfg = ctx->c;
  cudaStream_t st = (cudaStream_t)stream;
  const Ix x = make_ix(g);
  const int isc = x.isc, iec = x.iec, jsc = x.jsc, jec = x.jec, ied = x.ied, jed = x.jed, sj = g.sj, nz = g.nz;
  int rc;
  double *ucc = fv3::scratch_field(ctx, 16), *vcc = fv3::scratch_field(ctx, 17);
  double *fx = fv3::scratch_field(ctx, 18), *fy = fv3::scratch_field(ctx, 19);
  double *fx2 = fv3::scratch_field(ctx, 20), *fy2 = fv3::scratch_field(ctx, 21);
  double *dw = fv3::scratch_field(ctx, 22), *heat_s = fv3::scratch_field(ctx, 23);
  double *gxw = fv3::scratch_field(ctx, 24), *gyw = fv3::scratch_field(ctx, 25);
  double *gxq = fv3::scratch_field(ctx, 26), *gyq = fv3::scratch_field(ctx, 27);
  double *gxp = fv3::scratch_field(ctx, 28), *gyp = fv3::scratch_field(ctx, 29);
  double *ke = fv3::scratch_field(ctx, 30), *vort_a = fv3::scratch_field(ctx, 31);
  double *vort_b = fv3::scratch_field(ctx, 32), *abs_vort = fv3::scratch_field(ctx, 33);
  double *ut = fv3::scratch_field(ctx, 34), *vt = fv3::scratch_field(ctx, 35);
  const double *damp_w = c->damp_w, *ke_bg = c->ke_bg, *d_con = c->d_con, *damp_vt = c->damp_vt;

  // K1: contravariant winds, Courant numbers, area fluxes; cx += crx, cy += cry (fxadv.py:565-661, d_sw.py:29-50)
  if ((rc = fv3::fv_prep_launch(ctx, st, uc, vc, crx, cry, xfx, yfx, ucc, vcc, cx, cy, dt, false))) return rc;
  // K2: delp, w, q_con, pt transports and updates (d_sw.py:967-1090)
  double *side0 = fv3::scratch_field(ctx, 0);
  {
    auto mo = [](int hord) { const int a = hord < 0 ? -hord : hord; return a == 10 ? 8 : a; };
    const int mdp = mo(cfg.hord_dp), mvt = mo(cfg.hord_vt), mtm = mo(cfg.hord_tm);
#ifndef FV3_K2_SPLIT
#define FV3_K2_SPLIT 4
#endif
#define DSW_PART(A_, B_, C_, P_) \
  dsw_scalars_launch<A_, B_, C_, P_>(ctx, st, delp, pt, w, q_con, crx, cry, xfx, yfx, mfx, mfy, heat_s, diss_est, fx, fy, side0, dt, c, mdp, mvt, mtm)
#if FV3_K2_SPLIT == 4
#define DSW_SCALARS(A_, B_, C_) ((rc = DSW_PART(A_, B_, C_, 1)) || (rc = DSW_PART(A_, B_, C_, 2)) || (rc = DSW_PART(A_, B_, C_, 4)) || (rc = DSW_PART(A_, B_, C_, 8)), rc)
#elif FV3_K2_SPLIT == 2
#define DSW_SCALARS(A_, B_, C_) ((rc = DSW_PART(A_, B_, C_, 3)) || (rc = DSW_PART(A_, B_, C_, 12)), rc)
#else
#define DSW_SCALARS(A_, B_, C_) DSW_PART(A_, B_, C_, 15)
#endif
    if (mdp == 6 && mvt == 6 && mtm == 6)
      rc = DSW_SCALARS(6, 6, 6);
    else if (mdp == 5 && mvt == 5 && mtm == 5)
      rc = DSW_SCALARS(5, 5, 5);
    else if (mdp == 8 && mvt == 8 && mtm == 8)
      rc = DSW_SCALARS(8, 8, 8);
    else
      rc = DSW_SCALARS(0, 1, 2);  // mixed orders: one kernel with all three sweep bodies, selected per field
#undef DSW_SCALARS
#undef DSW_PART
    if (rc) return rc;
  }
  unpark_rows(ctx, st, Fields4{{w, q_con, pt, delp}}, 4, side0, nz);
  // K3a: kinetic energy, vorticity, divergence damping (d_sw.py:204-328, divergence_damping.py:482-632)
  const int mord = cfg.hord_mt < 0 ? -cfg.hord_mt : cfg.hord_mt;
  if (mord >= 8) {
    fv3::set_error("fv3_d_sw: hord_mt >= 8 is not implemented");
    return -1;
  }
  if (c->nonzero_nord > 3) {
    fv3::set_error("fv3_d_sw: nord > 3");
    return -1;
  }
  (void)delpc;  // scratch in the reference from here on; nothing downstream reads it
  if ((rc = dsw_vorticity_launch(ctx, st, u, v, ua, va, uc, vc, ucc, vcc, divgd, ke, vort_a, vort_b, dt, c))) return rc;
  // K3b: vorticity transport, wind update, vorticity damping and its heat (d_sw.py:1131-1237)
  {
    auto mo = [](int hord) { const int a = hord < 0 ? -hord : hord; return a == 10 ? 8 : a; };
    const int mvt = mo(cfg.hord_vt);
    double *side_u = fv3::scratch_field(ctx, 4);
#define DSW_WINDS(M_) dsw_winds_launch<M_>(ctx, st, u, v, ke, vort_a, vort_b, crx, cry, xfx, yfx, delp, heat_s, heat_source, side_u, dt, c)
    rc = mvt == 8 ? DSW_WINDS(8) : (mvt == 5 ? DSW_WINDS(5) : DSW_WINDS(6));
#undef DSW_WINDS
    if (rc) return rc;
    // the parked first u row of every strip but the first
    const fv3::StripGeom sg = fv3::strip_geometry(g, fv3::FVTP_PLANES, nz);
    if (sg.ns > 1) {
      const int R = sg.rows_per_strip;
      fv3::launch3d(ctx, st, isc, iec + 1, 0, sg.ns - 1, 0, nz, FV_LAMBDA(int s, int i, int bnd, int k) { FV_DEV_GM
        const int j = g.halo + (bnd + 1) * R;
        const int64_t o = O3(s, i, j, k);
        u[o] = side_u[o];
      });
    }
  }
  return fv3::check_launch("fv3_d_sw");
}

}  // extern "C"
