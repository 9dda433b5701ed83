// Non-hydrostatic pressure-gradient update of the D-grid winds and the remaining end-of-substep stencils.
//   fv3_nh_p_grad               <- NonHydrostaticPressureGradient.__call__ (nh_p_grad.py:190-255)
//   fv3_ray_fast                <- RayleighDamping.__call__ (ray_fast.py:184-206)
//   fv3_del2cubed               <- HyperdiffusionDamping.__call__ (del2cubed.py:165-194)
//   fv3_apply_diffusive_heating <- apply_diffusive_heating (temperature_adjust.py:8-43)
#include "a2b.h"
#include "common.h"
#include "plane.h"
#include "ppm.h"

extern "C" {

int fv3_nh_p_grad(fv3_ctx *ctx, double *u, double *v, double *pp, double *gz, double *pk3, const double *delp,
                  double dt, double ptop, double akap, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int h = g.halo, nz = g.nz, sj = g.sj;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const int64_t sk = g.sk;
  double *ppb = fv3::scratch_field(ctx, 16), *pk3b = fv3::scratch_field(ctx, 17), *gzb = fv3::scratch_field(ctx, 18);
  double *wk1 = fv3::scratch_field(ctx, 19);
  const double top_value = pow(ptop, akap);  // host libm, as `ptop ** akap` in the reference (:219)
  // four A->B interpolations (nh_p_grad.py:221-224) + set_k0 (:11-20): one plane-resident kernel per level
  // Five planes: the A-grid input is double-buffered — the next field's rows arrive by one bulk (TMA) copy while the
  // current field is interpolated, so no phase of the CTA waits on a load it has just issued.
  int rc = fv3::launch_planes(ctx, st, 0, nz + 1, 5, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *SQ0 = b.plane(0), *QX = b.plane(1), *QY = b.plane(2), *OUT = b.plane(3), *SQ1 = b.plane(4);
    const int64_t ob = O3(s, 0, 0, k);
    const int sj2 = g.sj, h2 = g.halo;
    auto fill = [&](double *dst, double value) {
      b.rect(h2, h2 + g.nx + 1, b.ja, b.jtop() + 1, [&](int i, int j) { dst[ob + j * sj2 + i] = value; });
    };
    b.prefetch_next_wave(gz, g, k);
    // fields of this interface level: gz always, pp and pk3 below the top interface, delp on layers
    const int nf = 1 + (k >= 1 ? 2 : 0) + (k < nz ? 1 : 0);
    auto field = [&](int f, const double *&src, double *&dst) {
      const int id = f == 0 ? 0 : (k >= 1 ? f : 3);  // 0 gz, 1 pp, 2 pk3, 3 delp
      src = id == 0 ? gz : id == 1 ? pp : id == 2 ? pk3 : delp;
      dst = id == 0 ? gzb : id == 1 ? ppb : id == 2 ? pk3b : wk1;
    };
    const double *src, *nxt;
    double *dst, *unused;
    field(0, src, dst);
    b.bulk_begin(1, sj2);
    b.bulk_rows(SQ0, src + ob, sj2);
#pragma unroll 1
    for (int f = 0; f < nf; ++f) {
      double *SQ = (f & 1) ? SQ1 : SQ0;
      field(f, src, dst);
      b.bulk_wait();
      if (f + 1 < nf) {
        field(f + 1, nxt, unused);
        b.bulk_begin(1, sj2);
        b.bulk_rows((f & 1) ? SQ0 : SQ1, nxt + ob, sj2);
      }
      fv3::a2b_plane(g, m, s, b, src + ob, SQ, QX, QY, OUT, true, dst + ob);  // B-grid values straight to the scratch field
    }
    if (k < 1) {
      fill(ppb, 0.0);
      fill(pk3b, top_value);
    }
  });
  if (rc) return rc;
  // replace pp, pk3, gz by their B-grid values; calc_u / calc_v (:23-112)
  fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, 0, nz + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    pp[o] = ppb[o];
    pk3[o] = pk3b[o];
    gz[o] = gzb[o];
    if (k == nz) return;
    if (i <= iec) {
      const double wk0 = pk3b[o + sk] - pk3b[o], wkx = pk3b[o + 1 + sk] - pk3b[o + 1];
      const double a = gzb[o + sk] - gzb[o + 1], b = gzb[o] - gzb[o + 1 + sk];
      const double du = dt / (wk0 + wkx) * (a * (pk3b[o + 1 + sk] - pk3b[o]) + b * (pk3b[o + sk] - pk3b[o + 1]));
      u[o] = (u[o] + du + dt / (wk1[o] + wk1[o + 1]) * (a * (ppb[o + 1 + sk] - ppb[o]) + b * (ppb[o + sk] - ppb[o + 1]))) * m.rdx[o2];
    }
    if (j <= jec) {
      const double wk0 = pk3b[o + sk] - pk3b[o], wky = pk3b[o + sj + sk] - pk3b[o + sj];
      const double a = gzb[o + sk] - gzb[o + sj], b = gzb[o] - gzb[o + sj + sk];
      const double dv = dt / (wk0 + wky) * (a * (pk3b[o + sj + sk] - pk3b[o]) + b * (pk3b[o + sk] - pk3b[o + sj]));
      v[o] = (v[o] + dv + dt / (wk1[o] + wk1[o + sj]) * (a * (ppb[o + sj + sk] - ppb[o]) + b * (ppb[o + sk] - ppb[o + sj]))) * m.rdy[o2];
    }
  });
  return fv3::check_launch("fv3_nh_p_grad");
}

// rf: per-level damping factor 1/(1+rfvals) (host-computed, ray_fast.py:22-38); n_rf = number of top levels with
// pfull < rf_cutoff, n_nudge = number with pfull < rf_cutoff_nudge, p_ref = sum of dp_ref over the latter.
int fv3_ray_fast(fv3_ctx *ctx, double *u, double *v, double *w, const double *rf, int n_rf, int n_nudge, double p_ref,
                 void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const double *dp = m.dp_ref;
  const int hydrostatic = ctx->c.hydrostatic;
  fv3::launch2d(ctx, (cudaStream_t)stream, isc, iec + 2, jsc, jec + 2, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    auto damp = [&](double *q) {
      double dm = 0.0;
      for (int k = 0; k < n_rf; ++k) {
        const double qv = q[c0 + k * sk];
        const double d = (1.0 - rf[k]) * dp[k] * qv;
        dm = k == 0 ? d : dm + d;
        q[c0 + k * sk] = qv * rf[k];
      }
      for (int k = 0; k < n_nudge; ++k) q[c0 + k * sk] = q[c0 + k * sk] + dm / p_ref;
    };
    if (i <= iec) damp(u);
    if (j <= jec) damp(v);
    if (!hydrostatic && i <= iec && j <= jec)
      for (int k = 0; k < n_rf; ++k) w[c0 + k * sk] = w[c0 + k * sk] * rf[k];
  });
  return fv3::check_launch("fv3_ray_fast");
}

int fv3_del2cubed(fv3_ctx *ctx, double *qdel, double cd, int nmax, int nk, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int h = g.halo;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const int ntimes = nmax < 3 ? nmax : 3;
  double *tmp = fv3::scratch_field(ctx, 16);
  double *bufs[2] = {qdel, tmp};
  int cur = 0;
  for (int n = 0; n < ntimes; ++n) {
    const int nt = ntimes - (n + 1);
    const double *qo = bufs[cur];
    double *qn = bufs[1 - cur];
    fv3::launch3d(ctx, st, isc - nt, iec + 1 + nt, jsc - nt, jec + 1 + nt, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
      const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
      const double third = 1.0 / 3.0;
      auto q0 = [&](int ii, int jj) { return qo[O3(s, ii, jj, k)]; };
      // corner_fill (del2cubed.py:25-56)
      auto qf = [&](int ii, int jj) {
        if (W && S) {
          if (ii == isc && jj == jsc) return (q0(ii, jj) + q0(ii - 1, jj) + q0(ii, jj - 1)) * third;
          if (ii == isc - 1 && jj == jsc) return (q0(ii + 1, jj) + q0(ii, jj) + q0(ii + 1, jj - 1)) * third;
          if (ii == isc && jj == jsc - 1) return (q0(ii, jj + 1) + q0(ii - 1, jj + 1) + q0(ii, jj)) * third;
        }
        if (E && S) {
          if (ii == iec && jj == jsc) return (q0(ii, jj) + q0(ii + 1, jj) + q0(ii, jj - 1)) * third;
          if (ii == iec + 1 && jj == jsc) return (q0(ii - 1, jj) + q0(ii, jj) + q0(ii - 1, jj - 1)) * third;
          if (ii == iec && jj == jsc - 1) return (q0(ii, jj + 1) + q0(ii + 1, jj + 1) + q0(ii, jj)) * third;
        }
        if (E && N) {
          if (ii == iec && jj == jec) return (q0(ii, jj) + q0(ii + 1, jj) + q0(ii, jj + 1)) * third;
          if (ii == iec + 1 && jj == jec) return (q0(ii - 1, jj) + q0(ii, jj) + q0(ii - 1, jj + 1)) * third;
          if (ii == iec && jj == jec + 1) return (q0(ii, jj - 1) + q0(ii + 1, jj - 1) + q0(ii, jj)) * third;
        }
        if (W && N) {
          if (ii == isc && jj == jec) return (q0(ii, jj) + q0(ii - 1, jj) + q0(ii, jj + 1)) * third;
          if (ii == isc - 1 && jj == jec) return (q0(ii + 1, jj) + q0(ii, jj) + q0(ii + 1, jj + 1)) * third;
          if (ii == isc && jj == jec + 1) return (q0(ii, jj - 1) + q0(ii - 1, jj - 1) + q0(ii, jj)) * third;
        }
        return q0(ii, jj);
      };
      auto qx = [&](int ii, int jj) {
        if (nt > 0) fv3::corner_x(g, s, ii, jj);
        return qf(ii, jj);
      };
      auto qy = [&](int ii, int jj) {
        if (nt > 0) fv3::corner_y(g, s, ii, jj);
        return qf(ii, jj);
      };
      auto fx = [&](int ii, int jj) { return m.del6_v[O2(s, ii, jj)] * (qx(ii - 1, jj) - qx(ii, jj)); };
      auto fy = [&](int ii, int jj) { return m.del6_u[O2(s, ii, jj)] * (qy(ii, jj - 1) - qy(ii, jj)); };
      // away from the cube corners (two cells either way) no corner fill or corner copy is read: plain neighbours
      const bool near_corner = (W || E) && (S || N) && (i <= isc + 1 || i >= iec - 1) && (j <= jsc + 1 || j >= jec - 1);
      if (!near_corner) {
        const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
        const int sj = g.sj;
        const double qc = qo[o];
        const double fx0 = m.del6_v[o2] * (qo[o - 1] - qc), fx1 = m.del6_v[o2 + 1] * (qc - qo[o + 1]);
        const double fy0 = m.del6_u[o2] * (qo[o - sj] - qc), fy1 = m.del6_u[o2 + sj] * (qc - qo[o + sj]);
        qn[o] = qc + cd * m.rarea[o2] * (fx0 - fx1 + fy0 - fy1);
        return;
      }
      // the copy `qdel = q` keeps the y-corner-copied field (the last in-place copy) outside the compute domain
      const double base = qy(i, j);
      qn[O3(s, i, j, k)] = base + cd * m.rarea[O2(s, i, j)] * (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1));
    });
    cur = 1 - cur;
  }
  if (cur == 1) {
    fv3::launch3d(ctx, st, isc, iec + 1, jsc, jec + 1, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
      const int64_t o = O3(s, i, j, k);
      qdel[o] = tmp[o];
    });
  }
  return fv3::check_launch("fv3_del2cubed");
}

int fv3_apply_diffusive_heating(fv3_ctx *ctx, const double *delp, const double *delz, const double *cappa,
                                const double *heat_source, double *pt, double delt_time_factor, int nk, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  const double RDG = -287.05 / 9.80665, CV_AIR = 1004.6 - 287.05;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    const double cp = cappa[o];
    const double pkz = pow(RDG * delp[o] / delz[o] * pt[o], cp / (1.0 - cp));
    const double dtmp = heat_source[o] / (CV_AIR * delp[o]);
    const double lim = k == 0 ? delt_time_factor * 0.1 : (k == 1 ? delt_time_factor * 0.5 : delt_time_factor);
    const double deltmin = fv3::rsign(fv3::dmin(lim, fabs(dtmp)), dtmp);
    pt[o] = pt[o] + deltmin / pkz;
  });
  return fv3::check_launch("fv3_apply_diffusive_heating");
}

}  // extern "C"
