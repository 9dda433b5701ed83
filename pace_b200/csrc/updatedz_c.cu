// Geopotential height on the C grid and the C-grid pressure-gradient update.
//   fv3_update_dz_c <- UpdateGeopotentialHeightOnCGrid.__call__ (fv3core/pace/fv3core/stencils/updatedzc.py:167-207)
//   fv3_p_grad_c    <- p_grad_c_stencil (dyn_core.py:120-171, non-hydrostatic branch)
// The reference's two corner-filled full copies of gz (gz_x, gz_y) are replaced by index remapping at read time.
#include "common.h"

namespace {
constexpr double DZ_MIN = 2.0;

// index of the element whose value fill_corners_2cells_x would have put at (i, j) (corners.py:130-166)
FV_HD void corner_src_x(const fv3_geom &g, int s, int isc, int iec, int jsc, int jec, int &i, int &j) {
  const bool S = fv3::on_south(g, s) && j == jsc - 1, N = fv3::on_north(g, s) && j == jec + 1;
  if (!(S || N)) return;
  if (fv3::on_west(g, s) && i < isc && i >= isc - 2) {
    const int d = isc - i;
    i = isc - 1;
    j = S ? jsc - 1 + d : jec + 1 - d;
  } else if (fv3::on_east(g, s) && i > iec && i <= iec + 2) {
    const int d = i - iec;
    i = iec + 1;
    j = S ? jsc - 1 + d : jec + 1 - d;
  }
}
FV_HD void corner_src_y(const fv3_geom &g, int s, int isc, int iec, int jsc, int jec, int &i, int &j) {
  const bool W = fv3::on_west(g, s) && i == isc - 1, E = fv3::on_east(g, s) && i == iec + 1;
  if (!(W || E)) return;
  if (fv3::on_south(g, s) && j < jsc && j >= jsc - 2) {
    const int d = jsc - j;
    j = jsc - 1;
    i = W ? isc - 1 + d : iec + 1 - d;
  } else if (fv3::on_north(g, s) && j > jec && j <= jec + 2) {
    const int d = j - jec;
    j = jec + 1;
    i = W ? isc - 1 + d : iec + 1 - d;
  }
}
}  // namespace

extern "C" {

int fv3_update_dz_c(fv3_ctx *ctx, const double *zs, const double *ut, const double *vt, double *gz, double *ws,
                    double dt, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int h = g.halo, nz = g.nz, sj = g.sj;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  double *gzn = fv3::scratch_field(ctx, 0);
  const double *dp0 = m.dp_ref;
  fv3::launch3d(ctx, st, isc - 1, iec + 2, jsc - 1, jec + 2, 0, nz + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    // pressure-weighted interpolation of the layer fluxes to interface k (updatedzc.py:15-33,91-100)
    auto iface = [&](const double *vel, int64_t ocol) {
      if (k == 0) {
        const double ratio = dp0[0] / (dp0[0] + dp0[1]);
        return vel[ocol] + (vel[ocol] - vel[ocol + g.sk]) * ratio;
      } else if (k == nz) {
        const double ratio = dp0[nz - 1] / (dp0[nz - 2] + dp0[nz - 1]);
        const double a = vel[ocol + (int64_t)(nz - 1) * g.sk], b = vel[ocol + (int64_t)(nz - 2) * g.sk];
        return a + (a - b) * ratio;
      }
      const double int_ratio = 1.0 / (dp0[k - 1] + dp0[k]);
      return (dp0[k] * vel[ocol + (int64_t)(k - 1) * g.sk] + dp0[k - 1] * vel[ocol + (int64_t)k * g.sk]) * int_ratio;
    };
    auto gzx = [&](int ii, int jj) {
      corner_src_x(g, s, isc, iec, jsc, jec, ii, jj);
      return gz[O3(s, ii, jj, k)];
    };
    auto gzy = [&](int ii, int jj) {
      corner_src_y(g, s, isc, iec, jsc, jec, ii, jj);
      return gz[O3(s, ii, jj, k)];
    };
    const int64_t c0 = O3(s, i, j, 0);
    const double xfx0 = iface(ut, c0), xfx1 = iface(ut, c0 + 1);
    const double yfx0 = iface(vt, c0), yfx1 = iface(vt, c0 + sj);
    const double fx0 = xfx0 * (xfx0 > 0.0 ? gzx(i - 1, j) : gzx(i, j));
    const double fx1 = xfx1 * (xfx1 > 0.0 ? gzx(i, j) : gzx(i + 1, j));
    const double fy0 = yfx0 * (yfx0 > 0.0 ? gzy(i, j - 1) : gzy(i, j));
    const double fy1 = yfx1 * (yfx1 > 0.0 ? gzy(i, j) : gzy(i, j + 1));
    const double ar = m.area[O2(s, i, j)];
    const int64_t o = c0 + (int64_t)k * g.sk;
    gzn[o] = (gz[o] * ar + fx0 - fx1 + fy0 - fy1) / (ar + xfx0 - xfx1 + yfx0 - yfx1);
  });
  fv3::launch2d(ctx, st, isc - 1, iec + 2, jsc - 1, jec + 2, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0);
    const double rdt = 1.0 / dt;
    double below = gzn[c0 + (int64_t)nz * g.sk];
    gz[c0 + (int64_t)nz * g.sk] = below;
    ws[O2(s, i, j)] = (zs[O2(s, i, j)] - below) * rdt;
    // the column walk is a dependent max-chain; its loads are not: 8 levels are fetched per trip before the chain runs
    // (with few subdomains per GPU there are too few columns to hide an L2 latency per level)
    int k = nz - 1;
    for (; k - 7 >= 0; k -= 8) {
      double v[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) v[n] = gzn[c0 + (int64_t)(k - n) * g.sk];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const double kp1 = below + DZ_MIN;
        below = v[n] > kp1 ? v[n] : kp1;
        gz[c0 + (int64_t)(k - n) * g.sk] = below;
      }
    }
    for (; k >= 0; --k) {
      const double v = gzn[c0 + (int64_t)k * g.sk], kp1 = below + DZ_MIN;
      below = v > kp1 ? v : kp1;
      gz[c0 + (int64_t)k * g.sk] = below;
    }
  });
  return fv3::check_launch("fv3_update_dz_c");
}

int fv3_p_grad_c(fv3_ctx *ctx, const double *rdxc, const double *rdyc, double *uc, double *vc, const double *delpc,
                 const double *pkc, const double *gz, double dt2, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, sj = g.sj;
  const int64_t sk = g.sk;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx + 1, h, h + g.ny + 1, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    uc[o] = uc[o] + dt2 * rdxc[o2] / (delpc[o - 1] + delpc[o]) *
                        ((gz[o - 1 + sk] - gz[o]) * (pkc[o + sk] - pkc[o - 1]) + (gz[o - 1] - gz[o + sk]) * (pkc[o - 1 + sk] - pkc[o]));
    vc[o] = vc[o] + dt2 * rdyc[o2] / (delpc[o - sj] + delpc[o]) *
                        ((gz[o - sj + sk] - gz[o]) * (pkc[o + sk] - pkc[o - sj]) + (gz[o - sj] - gz[o + sk]) * (pkc[o - sj + sk] - pkc[o]));
  });
  return fv3::check_launch("fv3_p_grad_c");
}

// gz_from_surface_height_and_thicknesses (dyn_core.py:83-96): compute domain
int fv3_gz_from_delz(fv3_ctx *ctx, const double *zs, const double *delz, double *gz, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, nz = g.nz;
  fv3::launch2d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0);
    double v = zs[O2(s, i, j)];
    gz[c0 + (int64_t)nz * g.sk] = v;
    int k = nz - 1;
    for (; k - 7 >= 0; k -= 8) {
      double d[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) d[n] = delz[c0 + (int64_t)(k - n) * g.sk];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        v = v - d[n];
        gz[c0 + (int64_t)(k - n) * g.sk] = v;
      }
    }
    for (; k >= 0; --k) {
      v = v - delz[c0 + (int64_t)k * g.sk];
      gz[c0 + (int64_t)k * g.sk] = v;
    }
  });
  return fv3::check_launch("fv3_gz_from_delz");
}

// interface_pressure_from_toa_pressure_and_thickness (dyn_core.py:99-112): compute domain + 1
int fv3_pem_from_delp(fv3_ctx *ctx, const double *delp, double *pem, double ptop, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, nz = g.nz;
  fv3::launch2d(ctx, (cudaStream_t)stream, h - 1, h + g.nx + 1, h - 1, h + g.ny + 1, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0);
    double v = ptop;
    pem[c0] = v;
    int k = 1;  // the reference stencil runs on nz levels and adds delp of the SAME level
    for (; k + 8 <= nz; k += 8) {
      double d[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) d[n] = delp[c0 + (int64_t)(k + n) * g.sk];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        v = v + d[n];
        pem[c0 + (int64_t)(k + n) * g.sk] = v;
      }
    }
    for (; k < nz; ++k) {
      v = v + delp[c0 + (int64_t)k * g.sk];
      pem[c0 + (int64_t)k * g.sk] = v;
    }
  });
  return fv3::check_launch("fv3_pem_from_delp");
}

// compute_geopotential (dyn_core.py:115-117): compute domain + 2, nz+1 levels
int fv3_compute_geopotential(fv3_ctx *ctx, const double *zh, double *gz, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  fv3::launch3d(ctx, (cudaStream_t)stream, h - 2, h + g.nx + 2, h - 2, h + g.ny + 2, 0, g.nz + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    gz[o] = zh[o] * 9.80665;
  });
  return fv3::check_launch("fv3_compute_geopotential");
}

}  // extern "C"
