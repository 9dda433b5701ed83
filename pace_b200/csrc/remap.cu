// Lagrangian-to-Eulerian vertical remapping.
//   fv3_remap_prep      <- init_pe + moist_cv_pt_pressure + pn2_pk_delp (remapping.py:34-193)
//   fv3_map_multi       <- n independent MapSingle.__call__ (map_single.py:147-200) in one launch: RemapProfile
//                          (remap_profile.py:150-563, kord 9) + lagrangian_contributions (map_single.py:21-81);
//                          MapNTracer (mapn_tracer.py:60-82) is the same call over the tracers
//   fv3_map_single      <- one MapSingle.__call__ (a batch of one)
//   fv3_fillz           <- FillNegativeTracerValues.__call__ (fillz.py:15-163)
//   fv3_remap_post      <- undo_delz_adjust_and_copy_peln + moist_pkz (remapping.py:46-59, moist_cv.py:112-141)
//   fv3_remap_pressures <- pressures_mapu / pressures_mapv (remapping.py:196-254)
//   fv3_remap_finish    <- update_ua + copy_from_below + moist_pt_last_step / adjust_divide (remapping.py:257-272,
//                          674-695)
//   fv3_fv_setup        <- fv_setup + pt_to_potential_density_pt (moist_cv.py:175-234, fv_dynamics.py:39-52)
#include "common.h"
#include "stream.h"

namespace {

constexpr int NKMAX = 96;
constexpr double RDGAS = 287.05, GRAV = 9.80665, CP_AIR = 1004.6, RVGAS = 461.50;
constexpr double RDG = -RDGAS / GRAV, CV_AIR = CP_AIR - RDGAS, CV_VAP = 3.0 * RVGAS, C_LIQ = 4.1855e3, C_ICE = 1972.0;
constexpr double ZVIR = RVGAS / RDGAS - 1;

FV_HD void posdef_iv1(double &a1, double &a2, double &a3, double &a4) {
  const double da1 = a3 - a2, da2 = da1 * da1, a6da = a4 * da1;
  if (((a1 - a2) * (a1 - a3)) >= 0.0) {
    a2 = a1;
    a3 = a1;
    a4 = 0.0;
  } else if (a6da < -1.0 * da2) {
    a4 = 3.0 * (a2 - a1);
    a3 = a2 - a4;
  } else if (a6da > da2) {
    a4 = 3.0 * (a3 - a1);
    a2 = a3 - a4;
  }
}
FV_HD void remap_constraint(double &a1, double &a2, double &a3, double &a4, bool extm) {
  const double da1 = a3 - a2, da2 = da1 * da1, a6da = a4 * da1;
  if (extm) {
    a2 = a1;
    a3 = a1;
    a4 = 0.0;
  } else if (a6da < -da2) {
    a4 = 3.0 * (a2 - a1);
    a3 = a2 - a4;
  } else if (a6da > da2) {
    a4 = 3.0 * (a3 - a1);
    a2 = a3 - a4;
  }
}
FV_HD void posdef_iv0(double &a1, double &a2, double &a3, double &a4) {
  if (a1 <= 0.0) {
    a2 = a1;
    a3 = a1;
    a4 = 0.0;
  } else if (fabs(a3 - a2) < -a4 && (a1 + 0.25 * ((a3 - a2) * (a3 - a2)) / a4 + a4 * (1.0 / 12.0)) < 0.0) {
    if ((a1 < a3) && (a1 < a2)) {
      a3 = a1;
      a2 = a1;
      a4 = 0.0;
    } else if (a3 > a2) {
      a4 = 3.0 * (a2 - a1);
      a3 = a2 - a4;
    } else {
      a4 = 3.0 * (a3 - a1);
      a2 = a3 - a4;
    }
  }
}
FV_HD double min3(double a, double b, double c) { return (a < b && a < c) ? a : (b < c ? b : c); }
FV_HD double max3(double a, double b, double c) { return (a > b && a > c) ? a : (b > c ? b : c); }

FV_HD void moist_cv(double qv, double ql_, double qr, double qs_, double qi, double qg, double &cvm, double &gz) {
  const double ql = ql_ + qr, qs = qi + qs_ + qg;
  gz = ql + qs;
  cvm = (1.0 - (qv + gz)) * CV_AIR + qv * CV_VAP + ql * C_LIQ + qs * C_ICE;
}

// ---- MapSingle as independent "chains" (one column of one field) -----------------------------------------------------
// A chain is one thread; its per-level arrays — the interface values q[0..km], the tridiagonal factors gam[k] and later
// the remapped column — are private columns of two scratch fields ([level][j][i]: coalesced across the lanes of a warp,
// L2-resident between the passes of the chain), so that occupancy is bounded by registers only (a first version kept q
// in shared memory, 20 KB per warp = 11 warps per SM: 2.25 ms per call against 2.0 ms now).  The layer means a1 and the
// pressures are re-read from the (L2-resident) inputs, and the PPM coefficients of a source layer are rebuilt from q and
// a1 when the Lagrangian walk enters it.  All fields of a remap call go in one launch.
// Expressions and their order follow the reference statement by statement (RemapProfile.__call__ for kord 9,
// remap_profile.py:622-681).
constexpr int MAP_MAX = 16;
constexpr int MAP_NT = 32;
struct MapBatch {
  int n;
  double *q1[MAP_MAX];
  const double *pe1[MAP_MAX], *pe2[MAP_MAX], *qs[MAP_MAX];
  double *out[MAP_MAX];
  double qmin[MAP_MAX];
  int qs2d[MAP_MAX], iv[MAP_MAX], iex[MAP_MAX], jex[MAP_MAX];
  int64_t qoff;  // the interface values of a chain live 16 scratch fields after its output column
};

// set_interpolation_coefficients (remap_profile.py:340-563) of source layer L from its interface values (b2, b3 on
// entry) and the layer means a1[L-2 .. L+2] = (am2, am1, v1, ap1, ap2)
FV_HD void map_coef(int L, int km, int iv, double qmin, double am2, double am1, double v1, double ap1, double ap2,
                    double &b1, double &b2, double &b3, double &b4) {
  b1 = v1;
  if (L == 0) {
    if (iv == 0 && b2 < 0.0) b2 = 0.0;
    if (iv == -1 && b2 * v1 <= 0.0) b2 = 0.0;
    b4 = 3.0 * (2.0 * v1 - (b2 + b3));
    posdef_iv1(b1, b2, b3, b4);
    return;
  }
  if (L == km - 1) {
    if (iv == 0 && b3 < 0.0) b3 = 0.0;
    if (iv == -1 && b3 * v1 <= 0.0) b3 = 0.0;
    b4 = 3.0 * (2.0 * v1 - (b2 + b3));
    posdef_iv1(b1, b2, b3, b4);
    return;
  }
  const double g0 = v1 - am1, g1 = ap1 - v1;  // gam[L], gam[L+1] (differences of the layer means)
  const bool ext0 = g0 * g1 < 0.0;
  if (L == 1 || L == km - 2) {
    b4 = 3.0 * (2.0 * v1 - (b2 + b3));
    remap_constraint(b1, b2, b3, b4, ext0);
    return;
  }
  const double gm1 = am1 - am2, g2 = ap2 - ap1;  // gam[L-1], gam[L+2]
  const bool extm = gm1 * g0 < 0.0, extp = g1 * g2 < 0.0;
  const double pmp_1 = v1 - 2.0 * g1, lac_1 = pmp_1 + 1.5 * g2;
  const double pmp_2 = v1 + 2.0 * g0, lac_2 = pmp_2 - 1.5 * gm1;
  if ((ext0 && extm) || (ext0 && extp) || (ext0 && (qmin > 0.0 && v1 < qmin))) {
    b2 = v1;
    b3 = v1;
    b4 = 0.0;
  } else {
    b4 = 6.0 * v1 - 3.0 * (b2 + b3);
    if (fabs(b4) > fabs(b2 - b3)) {
      double tmin = min3(v1, pmp_1, lac_1), tmax = max3(v1, pmp_1, lac_1);
      double t0 = b2 > tmin ? b2 : tmin;
      b2 = t0 < tmax ? t0 : tmax;
      tmin = min3(v1, pmp_2, lac_2);
      tmax = max3(v1, pmp_2, lac_2);
      t0 = b3 > tmin ? b3 : tmin;
      b3 = t0 < tmax ? t0 : tmax;
      b4 = 6.0 * v1 - 3.0 * (b2 + b3);
    }
  }
  if (iv == 0) posdef_iv0(b1, b2, b3, b4);
}


// Inputs are read through the read-only path (FV_LDG) and streamed a few levels ahead of their use: q1 is only
// written by this chain's final copy, after its last read.
template <class QS, class GS>
FV_DEV void map_chain(const fv3_geom &g, const MapBatch &mb, int f, int s, int i, int j, int km, QS Q, GS G) {
  const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
  double *q1 = mb.q1[f];
  const double *pe1 = mb.pe1[f] + c0, *pe2 = mb.pe2[f] + c0;
  const double *a1p = q1 + c0;
  const int iv = mb.iv[f];
  const double qmin = mb.qmin[f];
  auto A1 = [&](int k) { return FV_LDG(a1p + k * sk); };
  auto P1 = [&](int k) { return FV_LDG(pe1 + k * sk); };
#ifndef FV3_HOSTSIM
  // the whole column of layer means is read by four passes: pull it into L2 once, now
  for (int k = 8; k < km; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(a1p + k * sk));
#endif
  // set_initial_vals (remap_profile.py:150-250)
  if (iv != -2) {
    double p_k = P1(1), p_k1 = P1(2);
    double dpm = p_k - P1(0), dpk = p_k1 - p_k;
    double am = A1(0), ak = A1(1);
    double qm, gm;
    {
      const double gr = dpk / dpm, bet = gr * (gr + 0.5);
      qm = ((gr + gr) * (gr + 1.0) * am + ak) / bet;
      gm = (1.0 + gr * (gr + 1.5)) / bet;
      Q(0) = qm;
      G(0) = gm;
    }
    double d4 = 0.0;
    // trip k consumes pe1[k+2] and a1[k+1] (to set up trip k+1)
    fv3::stream_up(1, km, [&](int k) { return k + 2 <= km ? P1(k + 2) : 0.0; }, [&](int k) { return k + 1 < km ? A1(k + 1) : 0.0; },
              [&](int k, double p_next, double a_next) {
                d4 = dpm / dpk;
                const double bet = 2.0 + d4 + d4 - gm;
                qm = (3.0 * (am + d4 * ak) - qm) / bet;
                gm = d4 / bet;
                Q(k) = qm;
                G(k) = gm;
                if (k + 1 < km) {
                  dpm = dpk;
                  am = ak;
                  p_k = p_k1;
                  p_k1 = p_next;
                  dpk = p_k1 - p_k;
                  ak = a_next;
                }
              });
    {
      // d4 = delp[km-2] / delp[km-1], ak = a1[km-1], am = a1[km-2], qm = q[km-1], gm = gam[km-1]
      const double a_bot = 1.0 + d4 * (d4 + 1.5);
      Q(km) = (2.0 * d4 * (d4 + 1.0) * ak + am - a_bot * qm) / (d4 * (d4 + 0.5) - a_bot * gm);
    }
    double qn = Q(km);
    fv3::stream_down(0, km, [&](int k) { return G(k); }, [&](int k, double gk) {
      qn = Q(k) - gk * qn;
      Q(k) = qn;
    });
  } else {
    double qsv = 0.0;
    if (mb.qs[f] != nullptr) qsv = mb.qs2d[f] ? mb.qs[f][O2(s, i, j)] : mb.qs[f][c0];
    double p_k = P1(1), p_k1 = P1(2);
    double dpm = p_k - P1(0), dpk = p_k1 - p_k;  // delp[0], delp[1]
    double am = A1(0), ak = A1(1);
    const double q0 = 1.5 * am;
    Q(0) = q0;
    double gm = 0.5;  // gam[1]
    G(1) = gm;
    double grm = dpm / dpk;  // gr[1]
    double qm = (3.0 * (am + ak) - q0) / (2.0 + grm + grm - gm);
    Q(1) = qm;
    fv3::stream_up(2, km, [&](int k) { return P1(k + 1); }, [&](int k) { return A1(k); }, [&](int k, double p_new, double a_new) {
      dpm = dpk;
      am = ak;
      p_k = p_k1;
      p_k1 = p_new;
      dpk = p_k1 - p_k;  // delp[k]
      ak = a_new;
      const double old_bet = 2.0 + grm + grm - gm;
      gm = grm / old_bet;  // gam[k]
      G(k) = gm;
      const double grk = dpm / dpk;  // gr[k]
      if (k < km - 1) {
        const double bet = 2.0 + grk + grk - gm;
        qm = (3.0 * (am + ak) - qm) / bet;
        Q(k) = qm;
      } else {
        Q(km - 1) = (3.0 * (am + ak) - grk * qsv - qm) / (2.0 + grk + grk - gm);
      }
      grm = grk;
    });
    Q(km) = qsv;
    double qn = Q(km - 1);
    fv3::stream_down(0, km - 1, [&](int k) { return G(k + 1); }, [&](int k, double gk1) {
      qn = Q(k) - gk1 * qn;
      Q(k) = qn;
    });
  }
  // apply_constraints (:253-337) on the interior interfaces
  {
    double a_m2 = 0.0, a_m1 = A1(0), a_0 = A1(1);
    fv3::stream_up(1, km, [&](int k) { return k + 1 < km ? A1(k + 1) : 0.0; }, [&](int) { return 0.0; }, [&](int k, double a_p1, double) {
      const double tmp = a_m1 > a_0 ? a_m1 : a_0, tmp2 = a_m1 < a_0 ? a_m1 : a_0;
      double qk = Q(k);
      if (k == 1 || k == km - 1) {
        if (qk >= tmp) qk = tmp;
        if (qk <= tmp2) qk = tmp2;
      } else {
        const double g_m1 = a_m1 - a_m2, g_p1 = a_p1 - a_0;
        if (g_m1 * g_p1 > 0) {
          if (qk >= tmp) qk = tmp;
          if (qk <= tmp2) qk = tmp2;
        } else if (g_m1 > 0) {
          if (qk <= tmp2) qk = tmp2;
        } else {
          if (qk >= tmp) qk = tmp;
          if (iv == 0 && qk < 0.0) qk = 0.0;
        }
      }
      Q(k) = qk;
      a_m2 = a_m1;
      a_m1 = a_0;
      a_0 = a_p1;
    });
  }
  // lagrangian_contributions (map_single.py:21-81).  L = absolute source layer.  Sliding register windows around
  // it: layer means w[n] = a1[L-2+n] (n = 0..7) and pressures p[n] = pe1[L+n] (n = 0..4); the element entering a
  // window is requested three layers before the coefficients of its layer are built.
  int L = 0;
  double w0 = 0.0, w1 = 0.0, w2 = A1(0), w3 = A1(1), w4 = A1(2), w5 = A1(3), w6 = A1(4), w7 = A1(5);
  double p0 = P1(0), p1 = P1(1), p2 = P1(2), p3 = P1(3), p4 = P1(4);
  auto shift = [&]() {  // L -> L + 1
    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6; w6 = w7;
    w7 = L + 6 < km ? A1(L + 6) : 0.0;
    p0 = p1; p1 = p2; p2 = p3; p3 = p4;
    p4 = L + 5 <= km ? P1(L + 5) : 0.0;
    L = L + 1;
  };
  double b1, b2 = Q(0), b3 = Q(1), b4;
  map_coef(0, km, iv, qmin, w0, w1, w2, w3, w4, b1, b2, b3, b4);
  double top = FV_LDG(pe2);
  double bot_next = FV_LDG(pe2 + sk);
  for (int k = 0; k < km; ++k) {
    const double bot = bot_next;
    if (k + 2 <= km) bot_next = FV_LDG(pe2 + (k + 2) * sk);
    const double dpL = p1 - p0;
    const double pl = (top - p0) / dpL;
    double out;
    if (bot <= p1) {
      const double pr = (bot - p0) / dpL;
      out = b2 + 0.5 * (b4 + b3 - b2) * (pr + pl) - b4 * 1.0 / 3.0 * (pr * (pr + pl) + pl * pl);
    } else {
      double qsum = (p1 - top) * (b2 + 0.5 * (b4 + b3 - b2) * (1.0 + pl) - b4 * 1.0 / 3.0 * (1.0 + pl * (1.0 + pl)));
      // L = L + 1; while (L + 1 <= km and pe1[L+1] < bot): qsum += dp1[L] * a1[L]; L += 1; then L = min(L, km-1)
      bool more = L < km - 1;
      if (more) shift();
      while (more && p1 < bot) {
        qsum += (p1 - p0) * w2;
        more = L < km - 1;
        if (more) shift();
      }
      b2 = Q(L);
      b3 = Q(L + 1);
      map_coef(L, km, iv, qmin, w0, w1, w2, w3, w4, b1, b2, b3, b4);
      const double dpn = p1 - p0;
      const double dp = bot - p0, esl = dp / dpn;
      qsum += dp * (b2 + 0.5 * esl * (b3 - b2 + b4 * (1.0 - (2.0 / 3.0) * esl)));
      out = qsum / (bot - top);
    }
    G(k) = out;
    top = bot;
  }
  // copy the remapped column over the input, a few levels per trip (loads first)
  for (int k = 0; k < km; k += 8) {
    double t[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) t[n] = k + n < km ? G(k + n) : 0.0;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (k + n < km) q1[c0 + (k + n) * sk] = t[n];
  }
}

#ifndef FV3_HOSTSIM
__global__ void __launch_bounds__(MAP_NT) kmap(const MapBatch mb, int km) {
  const fv3_geom &g = c_g;
  // field fastest: the chains of one column tile (which share pe1 / pe2) are resident together
  const int f = (int)blockIdx.x, s = (int)blockIdx.z;
  const int ni = g.nx + mb.iex[f], nj = g.ny + mb.jex[f];
  const int idx = (int)blockIdx.y * MAP_NT + (int)threadIdx.x;
  if (idx >= ni * nj) return;
  const int jr = idx / ni, i = g.halo + (idx - jr * ni), j = g.halo + jr;
  double *gs = mb.out[f] + O3(s, i, j, 0);
  const int64_t sk = g.sk;
  double *qg = gs + mb.qoff;
  map_chain(g, mb, f, s, i, j, km, [&](int k) -> double & { return qg[k * sk]; },
            [&](int k) -> double & { return gs[k * sk]; });
}
#endif

}  // namespace

extern "C" {

int fv3_map_multi(fv3_ctx *ctx, int n, const int64_t *desc, const double *qmin, int kord, void *stream) {
  const fv3_geom g = ctx->g;
  if (kord < 0) kord = -kord;
  if (kord != 9) {
    fv3::set_error("fv3_map_multi: only kord 9 is implemented");
    return -1;
  }
  if (g.nz + 1 > NKMAX || g.nz < 8) {
    fv3::set_error("fv3_map_multi: nz out of range");
    return -1;
  }
  if (n < 1 || n > MAP_MAX) {
    fv3::set_error("fv3_map_multi: 1..16 fields per call");
    return -1;
  }
  MapBatch mb;
  mb.n = n;
  for (int f = 0; f < n; ++f) {
    const int64_t *d = desc + 8 * f;
    mb.q1[f] = (double *)d[0];
    mb.pe1[f] = (const double *)d[1];
    mb.pe2[f] = (const double *)d[2];
    mb.qs[f] = (const double *)d[3];
    mb.qs2d[f] = (int)d[4];
    mb.iv[f] = (int)d[5];
    mb.iex[f] = (int)d[6];
    mb.jex[f] = (int)d[7];
    mb.qmin[f] = qmin[f];
    mb.out[f] = fv3::scratch_field(ctx, f);
  }
  mb.qoff = fv3::scratch_field(ctx, 16) - fv3::scratch_field(ctx, 0);
  const int h = g.halo, km = g.nz;
#ifdef FV3_HOSTSIM
  (void)stream;
  for (int f = 0; f < n; ++f) {
    const int ni = g.nx + mb.iex[f], nj = g.ny + mb.jex[f];
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
    for (int s = 0; s < g.n_sub; ++s)
      for (int j = h; j < h + nj; ++j)
        for (int i = h; i < h + ni; ++i) {
          double qa[NKMAX], ga[NKMAX];
          map_chain(g, mb, f, s, i, j, km, [&](int k) -> double & { return qa[k]; },
                    [&](int k) -> double & { return ga[k]; });
        }
  }
  return 0;
#else
  cudaStream_t st = (cudaStream_t)stream;
  fv3::activate(ctx, st);
  const int ncols = (g.nx + 1) * (g.ny + 1);
  kmap<<<dim3(n, (ncols + MAP_NT - 1) / MAP_NT, g.n_sub), MAP_NT, 0, st>>>(mb, km);
  ++fv3::g_launches;
  return fv3::check_launch("fv3_map_multi");
#endif
}

int fv3_map_single(fv3_ctx *ctx, double *q1, const double *pe1, const double *pe2, const double *qs, int qs_is_2d,
                   double qmin, int kord, int iv, int i_extra, int j_extra, void *stream) {
  const int64_t d[8] = {(int64_t)q1, (int64_t)pe1, (int64_t)pe2, (int64_t)qs, qs_is_2d, iv, i_extra, j_extra};
  return fv3_map_multi(ctx, 1, d, &qmin, kord, stream);
}

int fv3_fillz(fv3_ctx *ctx, double *const *tracers, int nq, const double *dp2, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, km = g.nz;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, nq, FV_LAMBDA(int s, int i, int j, int t) { FV_DEV_GM
    double *qf = tracers[t];
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    {  // a column without negative values comes out unchanged (no fix-ups, zfix == 0): one streaming read, no write
      bool neg = false;
      for (int k = 0; k < km; ++k) neg = neg || qf[c0 + k * sk] < 0.0;
      if (!neg) return;
    }
    double q[NKMAX], dp[NKMAX], lower_fix[NKMAX], upper_fix[NKMAX], dm[NKMAX], dm_pos[NKMAX];
    bool any_neg = false;
    for (int k = 0; k < km; ++k) {
      q[k] = qf[c0 + k * sk];
      dp[k] = dp2[c0 + k * sk];
      lower_fix[k] = 0.0;
      upper_fix[k] = 0.0;
      any_neg = any_neg || q[k] < 0.0;
    }
    int zfix = 0;
    if (q[0] < 0.0) q[1] = q[1] + q[0] * dp[0] / dp[1];
    if (q[0] < 0) q[0] = 0;
    dm[0] = q[0] * dp[0];
    for (int k = 1; k < km - 1; ++k) {
      if (lower_fix[k - 1] != 0.0) q[k] = q[k] - (lower_fix[k - 1] / dp[k]);
      if (q[k] < 0.0) {
        zfix += 1;
        if (q[k - 1] > 0.0) {
          const double a = q[k - 1] * dp[k - 1], b = -(q[k] * dp[k]);
          const double dq = a < b ? a : b;
          q[k] = q[k] + dq / dp[k];
          upper_fix[k] = dq;
        }
        if ((q[k] < 0.0) && (q[k + 1] > 0.0)) {
          const double a = q[k + 1] * dp[k + 1], b = -(q[k] * dp[k]);
          const double dq = a < b ? a : b;
          q[k] = q[k] + dq / dp[k];
          lower_fix[k] = dq;
        }
      }
    }
    for (int k = 0; k < km - 1; ++k) {
      if (upper_fix[k + 1] != 0.0) q[k] = q[k] - upper_fix[k + 1] / dp[k];
      dm[k] = q[k] * dp[k];
      dm_pos[k] = dm[k] > 0.0 ? dm[k] : 0.0;
    }
    {
      const int k = km - 1;
      if (lower_fix[k - 1] != 0.0) q[k] = q[k] - (lower_fix[k - 1] / dp[k]);
      const double qup = q[k - 1] * dp[k - 1], qly = -q[k] * dp[k];
      const double dup = qup < qly ? qup : qly;
      if ((q[k] < 0.0) && (q[k - 1] > 0.0)) {
        zfix += 1;
        q[k] = q[k] + (dup / dp[k]);
        upper_fix[k] = dup;
      }
      dm[k] = q[k] * dp[k];
      dm_pos[k] = dm[k] > 0.0 ? dm[k] : 0.0;
    }
    {
      const int k = km - 2;
      if (upper_fix[k + 1] != 0.0) {
        q[k] = q[k] - (upper_fix[k + 1] / dp[k]);
        dm[k] = q[k] * dp[k];
        dm_pos[k] = dm[k] > 0.0 ? dm[k] : 0.0;
      }
    }
    double sum0 = 0.0, sum1 = 0.0;
    for (int k = 1; k < km; ++k) {
      sum0 += dm[k];
      sum1 += dm_pos[k];
    }
    const double fac = sum0 > 0.0 ? sum0 / sum1 : 0.0;
    if (zfix > 0 && fac > 0.0)
      for (int k = 1; k < km; ++k) {
        const double v = fac * dm[k] / dp[k];
        q[k] = v > 0.0 ? v : 0.0;
      }
    (void)any_neg;
    for (int k = 0; k < km; ++k) qf[c0 + k * sk] = q[k];
  });
  return fv3::check_launch("fv3_fillz");
}

// tracers6: qvapor, qliquid, qrain, qsnow, qice, qgraupel (device array of 6 pointers)
int fv3_remap_prep(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *pt, double *cappa, double *delp,
                   double *delz, const double *pe, double *pe1, double *pe2, double *dp2, double *ps, double *pn2,
                   const double *peln, double *pk, double ptop, double akap, double r_vir, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, km = g.nz;
  // every statement is local to a level once the surface pressure pe[km] of the column is known: one thread per
  // cell (the reference's k-loops carry no dependence here)
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny + 1, 0, km + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk, o = c0 + k * sk;
    // init_pe on the (nx, ny+1) domain
    const double pek = pe[o];
    pe1[o] = pek;
    if (k == 0) pe2[o] = ptop;
    if (k == km) pe2[o] = pek;
    if (j >= h + g.ny) return;
    const double psv = pe[c0 + km * sk];
    if (k == km) {
      ps[O2(s, i, j)] = psv;
      pn2[o] = peln[o];
      return;
    }
    {
      double cvm, gz;
      moist_cv(tracers6[0][o], tracers6[1][o], tracers6[2][o], tracers6[3][o], tracers6[4][o], tracers6[5][o], cvm, gz);
      q_con[o] = gz;
      const double cp = RDGAS / (RDGAS + cvm / (1.0 + r_vir * tracers6[0][o]));
      cappa[o] = cp;
      const double dp_old = delp[o], dz_old = delz[o], pt_old = pt[o];
      pt[o] = pt_old * exp(cp / (1.0 - cp) * log(RDG * dp_old / dz_old * pt_old));
      delz[o] = -dz_old / dp_old;
    }
    const double p2k = k == 0 ? ptop : m.ak[k] + m.bk[k] * psv;
    const double p2n = k + 1 == km ? psv : m.ak[k + 1] + m.bk[k + 1] * psv;
    if (k >= 1) pe2[o] = p2k;
    const double d = p2n - p2k;
    dp2[o] = d;
    delp[o] = d;
    const double l = log(p2k);
    pn2[o] = l;
    pk[o] = exp(akap * l);
  });
  return fv3::check_launch("fv3_remap_prep");
}

int fv3_remap_post(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *pkz, const double *pt, double *cappa,
                   const double *delp, double *delz, double *peln, double *pe0, const double *pn2, double r_vir,
                   void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, km = g.nz;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, km + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    pe0[o] = peln[o];
    peln[o] = pn2[o];
    if (k == km) return;
    const double dz = -delz[o] * delp[o];
    delz[o] = dz;
    double cvm, gz;
    moist_cv(tracers6[0][o], tracers6[1][o], tracers6[2][o], tracers6[3][o], tracers6[4][o], tracers6[5][o], cvm, gz);
    q_con[o] = gz;
    const double cp = RDGAS / (RDGAS + cvm / (1.0 + r_vir * tracers6[0][o]));
    cappa[o] = cp;
    pkz[o] = exp(cp * log(RDG * delp[o] / dz * pt[o]));
  });
  return fv3::check_launch("fv3_remap_post");
}

// dir 0: pressures_mapu on (nx, ny+1); dir 1: pressures_mapv on (nx+1, ny)
int fv3_remap_pressures(fv3_ctx *ctx, const double *pe, double *pe0, double *pe3, int dir, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, km = g.nz;
  const int64_t off = dir == 0 ? g.sj : 1;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx + (dir == 1), h, h + g.ny + (dir == 0), 0, km + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), ob = O3(s, i, j, km);
    const double bkh = 0.5 * m.bk[k];
    if (k == 0) {
      pe0[o] = pe[o];
      pe3[o] = dir == 0 ? m.ak[0] + bkh * (pe[ob - off] + pe[ob]) : m.ak[0];
    } else {
      pe0[o] = 0.5 * (pe[o - off] + pe[o]);
      pe3[o] = m.ak[k] + bkh * (pe[ob - off] + pe[ob]);
    }
  });
  return fv3::check_launch("fv3_remap_pressures");
}

int fv3_remap_finish(fv3_ctx *ctx, double *const *tracers6, const double *pe2, double *pe, double *pt, const double *pkz,
                     int last_step, double dtmp, double r_vir, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, km = g.nz;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, km, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    if (k >= 1) pe[o] = pe2[o];
    if (last_step) {
      const double gz = tracers6[1][o] + tracers6[2][o] + tracers6[4][o] + tracers6[3][o] + tracers6[5][o];
      pt[o] = (pt[o] + dtmp * pkz[o]) / ((1.0 + r_vir * tracers6[0][o]) * (1.0 - gz));
    } else {
      pt[o] = pt[o] / pkz[o];
    }
  });
  return fv3::check_launch("fv3_remap_finish");
}

// compute_preamble (fv_dynamics.py:440-483): fv_setup + pt_to_potential_density_pt
int fv3_fv_setup(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *cvm_out, double *pkz, double *pt,
                 double *cappa, const double *delp, const double *delz, double *dp1, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    double cvm, qc;
    moist_cv(tracers6[0][o], tracers6[1][o], tracers6[2][o], tracers6[3][o], tracers6[4][o], tracers6[5][o], cvm, qc);
    cvm_out[o] = cvm;
    q_con[o] = qc;
    const double d1 = ZVIR * tracers6[0][o];
    dp1[o] = d1;
    const double cp = RDGAS / (RDGAS + cvm / (1.0 + d1));
    cappa[o] = cp;
    const double pz = exp(cp * log(RDG * delp[o] * pt[o] * (1.0 + d1) * (1.0 - qc) / delz[o]));
    pkz[o] = pz;
    pt[o] = pt[o] * (1.0 + d1) * (1.0 - qc) / pz;
  });
  return fv3::check_launch("fv3_fv_setup");
}

// omega_from_w (fv_dynamics.py:55-64)
int fv3_omega_from_w(fv3_ctx *ctx, const double *delp, const double *delz, const double *w, double *omga, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    omga[o] = delp[o] / delz[o] * w[o];
  });
  return fv3::check_launch("fv3_omega_from_w");
}

}  // extern "C"
