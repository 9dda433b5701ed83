// Contravariant C-grid winds, Courant numbers and area fluxes for the D-grid transport.
//   fv3_fv_prep <- FiniteVolumeFluxPrep.__call__ (fv3core/pace/fv3core/stencils/fxadv.py:565-661)
// The reference's 8 stencils with save/restore temporaries are regrouped into 4 launches whose read and write
// sets were checked to be disjoint (DESIGN.md "fxadv"): KA main + tile-edge division, KB edge-row/column
// averages, KC the eight corner 2x2 solves, KD fluxes.  Points the reference leaves holding stale data from
// the previous call (uc_contra in the two rows next to a south/north tile edge for i outside [isc, iec+1]) are
// not consumed by anything downstream; here they receive the main formula.
#include "common.h"

namespace {
FV_HD double contra(double v1, double v2, double cosa, double rsin2) { return (v1 - v2 * cosa) * rsin2; }
}  // namespace

extern "C" {

int fv3_fv_prep(fv3_ctx *ctx, const double *uc, const double *vc, double *crx, double *cry, double *xfx, double *yfx,
                double *ucc, double *vcc, double dt, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int h = g.halo, nz = g.nz, sj = g.sj;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const int ied = iec + h, jed = jec + h;

  // KA: main_uc_vc_contra + uc_contra_y_edge + vc_contra_x_edge (fxadv.py:9-58,93-104)
  fv3::launch3d(ctx, st, 0, ied + 1, 0, jed + 1, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    if (i >= isc - 1 && i <= iec + 2) {
      double r;
      if ((W && i == isc) || (E && i == iec + 1)) {
        const double a = uc[o];
        r = a > 0 ? a / m.sin_sg3[o2 - 1] : a / m.sin_sg1[o2];
      } else {
        const double v = 0.25 * (vc[o - 1] + vc[o] + vc[o - 1 + sj] + vc[o + sj]);
        r = contra(uc[o], v, m.cosa_u[o2], m.rsin_u[o2]);
      }
      ucc[o] = r;
    }
    if (j >= jsc - 1 && j <= jec + 2) {
      double r;
      if ((S && j == jsc) || (N && j == jec + 1)) {
        const double a = vc[o];
        r = a > 0 ? a / m.sin_sg4[o2 - sj] : a / m.sin_sg2[o2];
      } else {
        const double u = 0.25 * (uc[o - sj] + uc[o + 1 - sj] + uc[o] + uc[o + 1]);
        r = contra(vc[o], u, m.cosa_v[o2], m.rsin_v[o2]);
      }
      vcc[o] = r;
    }
  });

  // KB: vc_contra_y_edge (:61-90) on the two columns next to a west/east tile edge and
  //     uc_contra_x_edge (:107-133) on the two rows next to a south/north tile edge
  fv3::launch3d(ctx, st, isc - 1, iec + 2, jsc - 1, jec + 2, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    const bool icol = (W && (i == isc - 1 || i == isc)) || (E && (i == iec || i == iec + 1));
    const bool jrow = (S && (j == jsc - 1 || j == jsc)) || (N && (j == jec || j == jec + 1));
    if (icol && j >= jsc && j <= jec + 1) {
      const bool restored = (S && (j == jsc || j == jsc + 1)) || (N && (j == jec || j == jec + 1));
      if (!restored) {
        const double u = 0.25 * (ucc[o - sj] + ucc[o + 1 - sj] + ucc[o] + ucc[o + 1]);
        vcc[o] = contra(vc[o], u, m.cosa_v[o2], 1.0);
      }
    }
    if (jrow && i >= isc && i <= iec + 1) {
      const bool restored = (W && (i == isc || i == isc + 1)) || (E && (i == iec || i == iec + 1));
      if (!restored) {
        const double v = 0.25 * (vcc[o - 1] + vcc[o] + vcc[o - 1 + sj] + vcc[o + sj]);
        ucc[o] = contra(uc[o], v, m.cosa_u[o2], 1.0);
      }
    }
  });

  // KC: uc_contra_corners (:136-243) and vc_contra_corners (:246-352): 16 points per subdomain and level
  fv3::launch3d(ctx, st, 0, 16, 0, 1, 0, nz, FV_LAMBDA(int s, int id, int unused, int k) { FV_DEV_GM
    (void)unused;
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int which = id / 8;   // 0: uc_contra, 1: vc_contra
    const int c = id % 8;
    const double *cu = m.cosa_u, *cv = m.cosa_v;
    if (which == 0) {
      const bool west = (c & 1) == 0, south = (c & 2) == 0, first = (c & 4) == 0;
      if (!((west ? W : E) && (south ? S : N))) return;
      const int i = west ? isc + 1 : iec;
      const int ja = south ? jsc - 1 : jec;
      const int j = first ? ja : ja + 1;
      const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
      double r;
      if (west && first) {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2] * cv[o2 - 1]);
        r = (uc[o] - 0.25 * cu[o2] * (vcc[o - 1 + sj] + vcc[o + sj] + vcc[o] + vc[o - 1] -
                                     0.25 * cv[o2 - 1] * (ucc[o - 1] + ucc[o - 1 - sj] + ucc[o - sj]))) * damp;
      } else if (west) {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2] * cv[o2 - 1 + sj]);
        r = (uc[o] - 0.25 * cu[o2] * (vcc[o - 1] + vcc[o] + vcc[o + sj] + vc[o - 1 + sj] -
                                     0.25 * cv[o2 - 1 + sj] * (ucc[o - 1] + ucc[o - 1 + sj] + ucc[o + sj]))) * damp;
      } else if (first) {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2] * cv[o2]);
        r = (uc[o] - 0.25 * cu[o2] * (vcc[o + sj] + vcc[o - 1 + sj] + vcc[o - 1] + vc[o] -
                                     0.25 * cv[o2] * (ucc[o + 1] + ucc[o + 1 - sj] + ucc[o - sj]))) * damp;
      } else {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2] * cv[o2 + sj]);
        r = (uc[o] - 0.25 * cu[o2] * (vcc[o] + vcc[o - 1] + vcc[o - 1 + sj] + vc[o + sj] -
                                     0.25 * cv[o2 + sj] * (ucc[o + 1] + ucc[o + 1 + sj] + ucc[o + sj]))) * damp;
      }
      ucc[o] = r;
    } else {
      const bool west = (c & 1) == 0, south = (c & 2) == 0, first = (c & 4) == 0;
      if (!((west ? W : E) && (south ? S : N))) return;
      const int j = south ? jsc + 1 : jec;
      // "first" = the column on the low-i side of the tile edge (i_start-1 / i_end), else i_start / i_end+1
      const int ia = west ? isc - 1 : iec;
      const int i = first ? ia : ia + 1;
      const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
      double r;
      if (south && first) {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2 - sj] * cv[o2]);
        r = (vc[o] - 0.25 * cv[o2] * (ucc[o + 1 - sj] + ucc[o + 1] + ucc[o] + uc[o - sj] -
                                     0.25 * cu[o2 - sj] * (vcc[o - sj] + vcc[o - 1 - sj] + vcc[o - 1]))) * damp;
      } else if (south) {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2 + 1 - sj] * cv[o2]);
        r = (vc[o] - 0.25 * cv[o2] * (ucc[o - sj] + ucc[o] + ucc[o + 1] + uc[o + 1 - sj] -
                                     0.25 * cu[o2 + 1 - sj] * (vcc[o - sj] + vcc[o + 1 - sj] + vcc[o + 1]))) * damp;
      } else if (!first) {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2 + 1] * cv[o2]);
        r = (vc[o] - 0.25 * cv[o2] * (ucc[o] + ucc[o - sj] + ucc[o + 1 - sj] + uc[o + 1] -
                                     0.25 * cu[o2 + 1] * (vcc[o + sj] + vcc[o + 1 + sj] + vcc[o + 1]))) * damp;
      } else {
        const double damp = 1.0 / (1.0 - 0.0625 * cu[o2] * cv[o2]);
        r = (vc[o] - 0.25 * cv[o2] * (ucc[o + 1] + ucc[o + 1 - sj] + ucc[o - sj] + uc[o] -
                                     0.25 * cu[o2] * (vcc[o + sj] + vcc[o - 1 + sj] + vcc[o - 1]))) * damp;
      }
      vcc[o] = r;
    }
  });

  // KD: fxadv_fluxes_stencil (:355-390)
  fv3::launch3d(ctx, st, 0, ied + 1, 0, jed + 1, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    if (i >= isc && i <= iec + 1) {
      const double a = ucc[o];
      if (a > 0) {
        crx[o] = dt * a * m.rdxa[o2 - 1];
        xfx[o] = m.dy[o2] * dt * a * m.sin_sg3[o2 - 1];
      } else {
        crx[o] = dt * a * m.rdxa[o2];
        xfx[o] = m.dy[o2] * dt * a * m.sin_sg1[o2];
      }
    }
    if (j >= jsc && j <= jec + 1) {
      const double a = vcc[o];
      if (a > 0) {
        cry[o] = dt * a * m.rdya[o2 - sj];
        yfx[o] = m.dx[o2] * dt * a * m.sin_sg4[o2 - sj];
      } else {
        cry[o] = dt * a * m.rdya[o2];
        yfx[o] = m.dx[o2] * dt * a * m.sin_sg2[o2];
      }
    }
  });
  return fv3::check_launch("fv3_fv_prep");
}

}  // extern "C"
