// Contravariant C-grid winds, Courant numbers and area fluxes for the D-grid transport.
//   fv3_fv_prep <- FiniteVolumeFluxPrep.__call__ (fv3core/pace/fv3core/stencils/fxadv.py:565-661)
// ONE strip-resident kernel (plane.h): uc and vc are staged with two bulk (TMA) copies, the reference's 8 stencils run
// as 4 block-wide phases on shared planes (KA main + tile-edge division, KB edge-row / column averages, KC the eight
// corner 2x2 solves, KD Courant numbers and area fluxes), and when d_sw calls it the accumulation of the Courant
// numbers (flux_capacitor, d_sw.py:29-50: cx += crx, cy += cry) is applied in the same pass.  The read and write sets
// of KB were checked to be disjoint (DESIGN.md "fxadv").  Points the reference leaves holding stale data from the
// previous call (uc_contra in the two rows next to a south/north tile edge for i outside [isc, iec+1]) are not consumed
// by anything downstream; here they receive the main formula.
#include "common.h"
#include "plane.h"

namespace {
FV_HD double contra(double v1, double v2, double cosa, double rsin2) { return (v1 - v2 * cosa) * rsin2; }
}  // namespace

namespace fv3 {

// cx / cy: Courant-number accumulators (may be null).  store_all: write the contravariant winds everywhere (the
// stand-alone entry point); otherwise only the bands next to tile edges that the kinetic-energy stage of d_sw reads.
int fv_prep_launch(fv3_ctx *ctx, cudaStream_t st, const double *uc, const double *vc, double *crx, double *cry, double *xfx,
                   double *yfx, double *ucc_out, double *vcc_out, double *cx, double *cy, double dt, bool store_all) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  return launch_planes(ctx, st, 0, g.nz, 4, FV_LAMBDA(int s, int k, const Block &b) { FV_DEV_GM
    double *UC = b.plane(0), *VC = b.plane(1), *UCC = b.plane(2), *VCC = b.plane(3);
    const int h = g.halo, sj = g.sj;
    const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1, ied = iec + h, jed = jec + h;
    const bool W = on_west(g, s), E = on_east(g, s), S = on_south(g, s), N = on_north(g, s);
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const int r0 = b.r0, r1 = b.r1;
    b.bulk_begin(2, sj);
    b.bulk_rows(UC, uc + ob, sj);
    b.bulk_rows(VC, vc + ob, sj);
    b.bulk_wait();
    // 2-D metric terms of this subdomain (plane offsets p = j * sj + i)
    const double *sin_sg1 = m.sin_sg1 + o2b, *sin_sg2 = m.sin_sg2 + o2b, *sin_sg3 = m.sin_sg3 + o2b, *sin_sg4 = m.sin_sg4 + o2b;
    const double *cosa_u = m.cosa_u + o2b, *cosa_v = m.cosa_v + o2b, *rsin_u = m.rsin_u + o2b, *rsin_v = m.rsin_v + o2b;
    // KA: main_uc_vc_contra + uc_contra_y_edge + vc_contra_x_edge (fxadv.py:9-58,93-104).  uc_contra of row j reads vc of
    // rows j, j+1; vc_contra of row j reads uc of rows j-1, j: every row whose inputs are resident is evaluated.
    b.rect(0, ied + 1, r0, r1, [&](int i, int j) {
      const int p = j * sj + i;
      if (i >= isc - 1 && i <= iec + 2 && j <= jed && j + 1 < r1) {
        double r;
        if ((W && i == isc) || (E && i == iec + 1)) {
          const double a = UC[p];
          r = a > 0 ? a / sin_sg3[p - 1] : a / sin_sg1[p];
        } else {
          const double v = 0.25 * (VC[p - 1] + VC[p] + VC[p - 1 + sj] + VC[p + sj]);
          r = contra(UC[p], v, cosa_u[p], rsin_u[p]);
        }
        UCC[p] = r;
      }
      if (j >= jsc - 1 && j <= jec + 2 && j - 1 >= r0) {
        double r;
        if ((S && j == jsc) || (N && j == jec + 1)) {
          const double a = VC[p];
          r = a > 0 ? a / sin_sg4[p - sj] : a / sin_sg2[p];
        } else {
          const double u = 0.25 * (UC[p - sj] + UC[p + 1 - sj] + UC[p] + UC[p + 1]);
          r = contra(VC[p], u, cosa_v[p], rsin_v[p]);
        }
        VCC[p] = r;
      }
    });
    // KB: vc_contra_y_edge (:61-90) on the two columns next to a west/east tile edge and uc_contra_x_edge (:107-133) on
    // the two rows next to a south/north tile edge (tile-edge subdomains only)
    if (W || E || S || N) {
      b.rect(isc - 1, iec + 3, b.lo(jsc - 1, 2), b.hi(jec + 3, 2), [&](int i, int j) {
        const int p = j * sj + i;
        const bool icol = (W && (i == isc - 1 || i == isc)) || (E && (i == iec || i == iec + 1));
        const bool jrow = (S && (j == jsc - 1 || j == jsc)) || (N && (j == jec || j == jec + 1));
        if (icol && j >= jsc && j <= jec + 1) {
          const bool restored = (S && (j == jsc || j == jsc + 1)) || (N && (j == jec || j == jec + 1));
          if (!restored) {
            const double u = 0.25 * (UCC[p - sj] + UCC[p + 1 - sj] + UCC[p] + UCC[p + 1]);
            VCC[p] = contra(VC[p], u, cosa_v[p], 1.0);
          }
        }
        if (jrow && i >= isc && i <= iec + 1) {
          const bool restored = (W && (i == isc || i == isc + 1)) || (E && (i == iec || i == iec + 1));
          if (!restored) {
            const double v = 0.25 * (VCC[p - 1] + VCC[p] + VCC[p - 1 + sj] + VCC[p + sj]);
            UCC[p] = contra(UC[p], v, cosa_u[p], 1.0);
          }
        }
      });
    }
    // KC: uc_contra_corners (:136-243) and vc_contra_corners (:246-352): 16 points per plane, evaluated by every strip
    // whose resident rows cover the point's neighbourhood
    if ((W || E) && (S || N)) {
      b.par(16, [&](int id) {
        const int which = id / 8;  // 0: uc_contra, 1: vc_contra
        const int c = id % 8;
        const double *cu = cosa_u, *cv = cosa_v;
        const bool west = (c & 1) == 0, south = (c & 2) == 0, first = (c & 4) == 0;
        if (!((west ? W : E) && (south ? S : N))) return;
        if (which == 0) {
          const int i = west ? isc + 1 : iec;
          const int ja_ = south ? jsc - 1 : jec;
          const int j = first ? ja_ : ja_ + 1;
          if (j - 2 < r0 && r0 > 0) return;
          if (j + 2 >= r1 && r1 < g.nj) return;
          const int p = j * sj + i;
          double r;
          if (west && first) {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p] * cv[p - 1]);
            r = (UC[p] - 0.25 * cu[p] * (VCC[p - 1 + sj] + VCC[p + sj] + VCC[p] + VC[p - 1] -
                                        0.25 * cv[p - 1] * (UCC[p - 1] + UCC[p - 1 - sj] + UCC[p - sj]))) * damp;
          } else if (west) {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p] * cv[p - 1 + sj]);
            r = (UC[p] - 0.25 * cu[p] * (VCC[p - 1] + VCC[p] + VCC[p + sj] + VC[p - 1 + sj] -
                                        0.25 * cv[p - 1 + sj] * (UCC[p - 1] + UCC[p - 1 + sj] + UCC[p + sj]))) * damp;
          } else if (first) {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p] * cv[p]);
            r = (UC[p] - 0.25 * cu[p] * (VCC[p + sj] + VCC[p - 1 + sj] + VCC[p - 1] + VC[p] -
                                        0.25 * cv[p] * (UCC[p + 1] + UCC[p + 1 - sj] + UCC[p - sj]))) * damp;
          } else {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p] * cv[p + sj]);
            r = (UC[p] - 0.25 * cu[p] * (VCC[p] + VCC[p - 1] + VCC[p - 1 + sj] + VC[p + sj] -
                                        0.25 * cv[p + sj] * (UCC[p + 1] + UCC[p + 1 + sj] + UCC[p + sj]))) * damp;
          }
          UCC[p] = r;
        } else {
          const int j = south ? jsc + 1 : jec;
          // "first" = the column on the low-i side of the tile edge (i_start-1 / i_end), else i_start / i_end+1
          const int ia = west ? isc - 1 : iec;
          const int i = first ? ia : ia + 1;
          if (j - 2 < r0 && r0 > 0) return;
          if (j + 2 >= r1 && r1 < g.nj) return;
          const int p = j * sj + i;
          double r;
          if (south && first) {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p - sj] * cv[p]);
            r = (VC[p] - 0.25 * cv[p] * (UCC[p + 1 - sj] + UCC[p + 1] + UCC[p] + UC[p - sj] -
                                        0.25 * cu[p - sj] * (VCC[p - sj] + VCC[p - 1 - sj] + VCC[p - 1]))) * damp;
          } else if (south) {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p + 1 - sj] * cv[p]);
            r = (VC[p] - 0.25 * cv[p] * (UCC[p - sj] + UCC[p] + UCC[p + 1] + UC[p + 1 - sj] -
                                        0.25 * cu[p + 1 - sj] * (VCC[p - sj] + VCC[p + 1 - sj] + VCC[p + 1]))) * damp;
          } else if (!first) {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p + 1] * cv[p]);
            r = (VC[p] - 0.25 * cv[p] * (UCC[p] + UCC[p - sj] + UCC[p + 1 - sj] + UC[p + 1] -
                                        0.25 * cu[p + 1] * (VCC[p + sj] + VCC[p + 1 + sj] + VCC[p + 1]))) * damp;
          } else {
            const double damp = 1.0 / (1.0 - 0.0625 * cu[p] * cv[p]);
            r = (VC[p] - 0.25 * cv[p] * (UCC[p + 1] + UCC[p + 1 - sj] + UCC[p - sj] + UC[p] -
                                        0.25 * cu[p] * (VCC[p + sj] + VCC[p - 1 + sj] + VCC[p - 1]))) * damp;
          }
          VCC[p] = r;
        }
      });
    }
    // KD: fxadv_fluxes_stencil (:355-390) on the rows this strip owns (the first / last strip also owns the halo rows
    // below / above the compute domain), plus flux_capacitor's Courant-number accumulation
    const int jlo = b.first ? 0 : b.ja, jhi = b.last ? jed + 1 : b.jb;
    const double *rdxa = m.rdxa + o2b, *rdya = m.rdya + o2b, *dxm = m.dx + o2b, *dym = m.dy + o2b;
    b.rect(0, ied + 1, jlo, jhi, [&](int i, int j) {
      const int p = j * sj + i;
      const int64_t o = ob + p;
      if (i >= isc && i <= iec + 1) {
        const double a = UCC[p];
        double cr, xf;
        if (a > 0) {
          cr = dt * a * rdxa[p - 1];
          xf = dym[p] * dt * a * sin_sg3[p - 1];
        } else {
          cr = dt * a * rdxa[p];
          xf = dym[p] * dt * a * sin_sg1[p];
        }
        crx[o] = cr;
        xfx[o] = xf;
        if (cx) cx[o] = cx[o] + cr;
      }
      if (j >= jsc && j <= jec + 1 && (j < b.jb || b.last)) {
        const double a = VCC[p];
        double cr, yf;
        if (a > 0) {
          cr = dt * a * rdya[p - sj];
          yf = dxm[p] * dt * a * sin_sg4[p - sj];
        } else {
          cr = dt * a * rdya[p];
          yf = dxm[p] * dt * a * sin_sg2[p];
        }
        cry[o] = cr;
        yfx[o] = yf;
        if (cy) cy[o] = cy[o] + cr;
      }
      // contravariant winds for the kinetic-energy stage (d_sw.py:204-298 reads them next to tile edges only)
      const bool band = store_all || (S && j <= jsc + 2) || (N && j >= jec - 1) || (W && i <= isc + 2) || (E && i >= iec - 1);
      if (band) {
        if (i >= isc - 1 && i <= iec + 2 && j <= jed) ucc_out[o] = UCC[p];
        if (j >= jsc - 1 && j <= jec + 2) vcc_out[o] = VCC[p];
      }
    });
  });
}

}  // namespace fv3

extern "C" {

int fv3_fv_prep(fv3_ctx *ctx, const double *uc, const double *vc, double *crx, double *cry, double *xfx, double *yfx,
                double *ucc, double *vcc, double dt, void *stream) {
  int rc = fv3::fv_prep_launch(ctx, (cudaStream_t)stream, uc, vc, crx, cry, xfx, yfx, ucc, vcc, nullptr, nullptr, dt, true);
  if (rc) return rc;
  return fv3::check_launch("fv3_fv_prep");
}

}  // extern "C"
