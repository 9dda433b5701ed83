/*
 * fv3_b200.h — C ABI of the B200-native FV3 dynamical-core hot path.
 *
 * Every entry point replaces one call of the reference (ai2cm/pace; file:line given per function) and keeps
 * that call's argument order and meaning.  Rules of the boundary:
 *   - plain C: pointers, sizes, scalars; no C++/torch types; no exceptions cross it;
 *   - all array pointers are DEVICE pointers borrowed for the duration of the call (the Python side owns the
 *     memory through pace_b200.util.Quantity / __cuda_array_interface__); nothing is allocated here;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises;
 *   - return value 0 = ok, otherwise a cudaError_t (or -1 for argument errors); fv3_last_error() gives text.
 *
 * Array layout ("batched I-fastest"): a 3-D field holds all local subdomains,
 *     element (s, i, j, k) at  base[s*ss + k*sk + j*sj + i],   i, j include the halo (compute origin = halo),
 * with storage extents ni = nx+2*halo+1, nj = ny+2*halo+1, nk = nz+1 for EVERY field (as the reference's
 * SubtileGridSizer does, util/pace/util/initialization/sizer.py:142-155).  2-D fields use (s, i, j) at
 * base[s*ss2 + j*sj + i].  Column vectors (ak, bk, dp_ref, pfull ...) are plain double[nk].
 */
#ifndef FV3_B200_H
#define FV3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FV3_MAX_SUBDOMAINS 64
#define FV3_MAX_LEVELS 128
#define FV3_EDGE_WEST 1
#define FV3_EDGE_EAST 2
#define FV3_EDGE_SOUTH 4
#define FV3_EDGE_NORTH 8

typedef struct fv3_geom {
  int32_t n_sub;            /* local subdomains held by this process / GPU */
  int32_t nx, ny, nz;       /* compute cells per subdomain */
  int32_t halo;             /* 3 */
  int32_t ni, nj, nk;       /* storage extents */
  int32_t sj;               /* j stride (elements); i stride is 1 */
  int32_t pad_;
  int64_t sk, ss, ss2;      /* k stride, subdomain stride (3-D), subdomain stride (2-D) */
  uint8_t edge[FV3_MAX_SUBDOMAINS]; /* FV3_EDGE_* bits: is subdomain s on that edge of its cube tile
                                       (GridIndexing.west_edge..., dsl/pace/dsl/stencil.py:717-758) */
} fv3_geom;

/* Scalar namelist values read on the path (fv3core/pace/fv3core/_config.py:14-160). */
typedef struct fv3_config {
  int32_t hord_dp, hord_tm, hord_mt, hord_vt, hord_tr;
  int32_t kord_tm, kord_tr, kord_wz, kord_mt;
  int32_t nord, n_sponge, nwat, fill, do_vort_damp, convert_ke, hydrostatic, rf_fast;
  int32_t ks;               /* GridData.ks */
  double d2_bg, d2_bg_k1, d2_bg_k2, d4_bg, ke_bg, dddmp, vtdm4, d_con, delt_max;
  double p_fac, a_imp, tau, rf_cutoff;
  double ptop, da_min, da_min_c;
} fv3_config;

/* Device pointers to the 2-D metric terms of all local subdomains (util/pace/util/grid/helper.py:305-640)
 * and the damping coefficients (helper.py:20-60); columns are double[nk]. */
typedef struct fv3_grid {
  const double *dx, *dy, *dxa, *dya, *dxc, *dyc, *rdx, *rdy, *rdxa, *rdya, *rdxc, *rdyc;
  const double *area, *area_64, *rarea, *rarea_c;
  const double *cosa, *cosa_u, *cosa_v, *cosa_s, *sina_u, *sina_v, *rsina, *rsin_u, *rsin_v, *rsin2;
  const double *sin_sg1, *sin_sg2, *sin_sg3, *sin_sg4, *cos_sg1, *cos_sg2, *cos_sg3, *cos_sg4;
  const double *fC, *f0;
  const double *edge_w, *edge_e, *edge_s, *edge_n;       /* a2b_ord4 edge weights, stored as 2-D fields */
  const double *divg_u, *divg_v, *del6_u, *del6_v;
  const double *a11, *a12, *a21, *a22;
  const double *ak, *bk, *dp_ref, *pfull;                /* columns */
  const double *a2b_w;  /* [n_sub][4 corners sw,se,nw,ne][3 arms] great-circle extrapolation weights x1/(x2-x1)
                           of a2b_ord4.extrap_corner (a2b_ord4.py:37-56), computed once on the host */
} fv3_grid;

/* Per-level damping parameters of d_sw (get_column_namelist, d_sw.py:611-683) as device columns double[>= nz+1],
 * plus the del-n coefficient columns derived from them with calc_damp (delnflux.py:18-33) on the host. */
typedef struct fv3_dsw_cols {
  const double *nord, *nord_v, *nord_w, *nord_t, *damp_vt, *damp_w, *damp_t, *d_con, *ke_bg, *d2_divg;
  const double *dn_damp_vt;    /* calc_damp(damp_vt, da_min,   nord_v): DelnFlux inside fvtp2d_dp / fvtp2d_tm */
  const double *dn_damp_t;     /* calc_damp(damp_t,  da_min,   nord_t): DelnFlux inside fvtp2d_dp_t */
  const double *dn_damp_vt_c;  /* calc_damp(damp_vt, da_min_c, nord_v): delnflux_nosg_v (d_sw.py:915-919) */
  const double *dn_damp_w_c;   /* calc_damp(damp_w,  da_min_c, nord_w): delnflux_nosg_w (d_sw.py:920-924) */
  int32_t nmax_v, nmax_w, nmax_t;        /* max over levels of the nord columns */
  int32_t nonzero_nord_k, nonzero_nord;  /* first level with nord > 0 and its nord (divergence_damping.py:216-223) */
} fv3_dsw_cols;

/* SatAdjustConfig (fv3core/pace/fv3core/_config.py:16-40): the externals and time scales of the fast saturation
 * adjustment. */
typedef struct fv3_sat_adjust_config {
  int32_t hydrostatic, rad_snow, rad_rain, rad_graupel, tintqs, icloud_f;
  double sat_adj0, ql_gen, qs_mlt, ql0_max, t_sub, qi_gen, qi_lim, qi0_max, dw_ocean, dw_land, cld_min;
  double tau_i2s, tau_v2l, tau_r2g, tau_l2r, tau_l2v, tau_imlt, tau_smlt;
} fv3_sat_adjust_config;

typedef struct fv3_ctx fv3_ctx;

/* scratch: device buffer of scratch_bytes used for stage-private temporaries (never freed here). */
fv3_ctx *fv3_create(const fv3_geom *geom, const fv3_config *config, const fv3_grid *grid, void *scratch,
                    int64_t scratch_bytes);
void fv3_destroy(fv3_ctx *ctx);
const char *fv3_last_error(void);
int fv3_abi_version(void);
/* 1 when this library was built as the CPU host-simulation of the kernels (tests only), 0 for CUDA. */
int fv3_is_hostsim(void);
/* CUDA kernels launched by this library since it was loaded (bench.py reports the difference over the timed region) */
int64_t fv3_launch_count(void);
/* measurement aid: blocks x 128 threads run iters x 8 independent fp64 FMAs each (2 * 8 * iters * 128 * blocks flops);
 * out: blocks * 128 doubles.  bench.py times it to get the fp64 roof of this GPU. */
int fv3_fp64_peak(double *out, int iters, int blocks, void *stream);
/* ---- optional state sanity check: replaces the min / max / NaN scans of SafetyChecker.check_state
 *      (driver/pace/driver/safety_checks.py:70-110) and the negative-delp / negative-tracer / NaN checks of the DaCe
 *      debug passes (dsl/pace/dsl/dace/sdfg_debug_passes.py:185-269) by one pass over the field.
 * out (device, 3 x int64): key of the minimum and of the maximum over the non-NaN values — key = bits ^ ((bits >> 63) &
 * 0x7fffffffffffffff), an order-preserving involution of the fp64 bit pattern — and the number of NaN values, over the
 * points i in [i0, i1), j in [j0, j1), levels [0, nk) of every local subdomain (a Quantity's view or its whole storage);
 * nk == 0: 2-D field. */
int fv3_field_check(fv3_ctx *ctx, const double *field, int i0, int i1, int j0, int j1, int nk, int64_t *out,
                    void *stream);
/* number of 3-D scratch fields fv3_create needs (scratch_bytes >= n * ss * n_sub * 8) */
int fv3_scratch_fields(void);

/* ---- halo exchange: replaces HaloUpdater.start/wait pack/unpack + the four NVRTC kernels
 *      (util/pace/util/halo_updater.py:217-303, halo_data_transformer.py:537-921, cuda_kernels.py:5-178).
 * One launch moves n_entries 2-D points x nlev levels x n_fields fields.  Entry e copies
 *   dst = fields[dst_comp[e]*n_fields + f] + dst_off[e] (+ k*sk)  <-  sign[e] * (src likewise),
 * offsets already include the subdomain stride.  fields holds n_comp*n_fields device pointers.        */
int fv3_halo_gather(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *dst_off,
                    const int64_t *src_off, const int8_t *dst_comp, const int8_t *src_comp, const double *sign,
                    int64_t n_entries, void *stream);
/* pack: buf[(f*nlev + k)*n_entries + e] = sign[e] * src ;  unpack: dst = buf[...] (sign already applied) */
int fv3_halo_pack(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *src_off,
                  const int8_t *src_comp, const double *sign, int64_t n_entries, double *buf, void *stream);
int fv3_halo_unpack(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *dst_off,
                    const int8_t *dst_comp, int64_t n_entries, const double *buf, void *stream);
/* ---- inter-GPU messages of an exchange as one C call: replaces the Isend / Irecv pairs of
 *      HaloUpdater.start (util/pace/util/halo_updater.py:217-303) by ONE grouped ncclSend / ncclRecv per peer GPU on the
 *      packed segments below, asynchronous on `stream`.  NCCL is resolved at run time from the process's libnccl.so.2
 *      (fv3_nccl_available() == 0 when there is none).  A communicator is created collectively from 128 id bytes that
 *      rank 0 obtains and distributes (fv3_nccl_unique_id, fv3_nccl_comm_create); an integration that already owns an
 *      ncclComm_t passes it straight in.  Offsets and counts are in doubles, host arrays. */
int fv3_nccl_available(void);
int fv3_nccl_unique_id(char *id128);
int fv3_nccl_comm_create(void **comm, int nranks, const char *id128, int rank);
int fv3_nccl_comm_destroy(void *comm);
int fv3_halo_exchange_nccl(void *nccl_comm, const double *send_buf, const int64_t *send_off, const int64_t *send_cnt,
                           const int32_t *send_peer, int n_send, double *recv_buf, const int64_t *recv_off,
                           const int64_t *recv_cnt, const int32_t *recv_peer, int n_recv, void *stream);
/* segmented forms: ONE launch per exchange for all peer GPUs.  buf holds one contiguous segment per peer (the NCCL
 * send / recv message, halo_updater.py:217-303 posts one Isend/Irecv per neighbour); entry e lives at
 *   buf[seg_base[e] + (f*nlev + k)*seg_n[e] + seg_e[e]].                                                   */
int fv3_halo_pack_segments(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *src_off,
                           const int8_t *src_comp, const double *sign, const int64_t *seg_base, const int32_t *seg_n,
                           const int32_t *seg_e, int64_t n_entries, double *buf, void *stream);
int fv3_halo_unpack_segments(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *dst_off,
                             const int8_t *dst_comp, const int64_t *seg_base, const int32_t *seg_n, const int32_t *seg_e,
                             int64_t n_entries, const double *buf, void *stream);

/* ---- NonhydrostaticVerticalSolverCGrid.__call__ (fv3core/pace/fv3core/stencils/riem_solver_c.py:172-250) */
int fv3_riem_solver_c(fv3_ctx *ctx, double dt2, const double *cappa, double ptop, const double *hs,
                      const double *ws, const double *ptc, const double *q_con, const double *delpc, double *gz,
                      double *pef, const double *w3, void *stream);

/* ---- CGridShallowWaterDynamics.__call__ (fv3core/pace/fv3core/stencils/c_sw.py:607-766), including
 *      DGrid2AGrid2CGridVectors (d2a2c_vect.py:547-655).  delpc/ptc are the stage's own outputs
 *      (self.delpc, self.ptc in the reference, returned by __call__). */
int fv3_c_sw(fv3_ctx *ctx, double *delp, double *pt, const double *u, const double *v, double *w, double *uc,
             double *vc, double *ua, double *va, double *ut, double *vt, double *divgd, double *omga, double *delpc,
             double *ptc, double dt2, void *stream);

/* ---- UpdateGeopotentialHeightOnCGrid.__call__ (updatedzc.py:167-207); dp_ref and area come from fv3_grid */
int fv3_update_dz_c(fv3_ctx *ctx, const double *zs, const double *ut, const double *vt, double *gz, double *ws,
                    double dt, void *stream);
/* ---- p_grad_c_stencil (dyn_core.py:120-171), non-hydrostatic */
int fv3_p_grad_c(fv3_ctx *ctx, const double *rdxc, const double *rdyc, double *uc, double *vc, const double *delpc,
                 const double *pkc, const double *gz, double dt2, void *stream);
/* ---- small AcousticDynamics stencils (dyn_core.py:83-117) */
int fv3_gz_from_delz(fv3_ctx *ctx, const double *zs, const double *delz, double *gz, void *stream);
int fv3_pem_from_delp(fv3_ctx *ctx, const double *delp, double *pem, double ptop, void *stream);
int fv3_compute_geopotential(fv3_ctx *ctx, const double *zh, double *gz, void *stream);

/* ---- FiniteVolumeFluxPrep.__call__ (fv3core/pace/fv3core/stencils/fxadv.py:565-661) */
int fv3_fv_prep(fv3_ctx *ctx, const double *uc, const double *vc, double *crx, double *cry, double *xfx, double *yfx,
                double *uc_contra, double *vc_contra, double dt, void *stream);

/* ---- FiniteVolumeTransport.__call__ (fvtp2d.py:235-346).  Optional (may be NULL): x/y_mass_flux, mass,
 *      and the per-level columns nord_col / damp_col (device double[nk]) that switch on DelnFlux
 *      (delnflux.py:1164-1207; damp_col = calc_damp(damp_c, da_min, nord), delnflux.py:18-33; nmax = max nord).
 *      nk = number of levels (nz, or nz+1 for interface fields).  Valid outputs: fx on [isc..iec+1]x[jsc..jec],
 *      fy on [isc..iec]x[jsc..jec+1].  q is not modified (the reference rewrites q's cube-corner halo cells). */
int fv3_fvtp2d(fv3_ctx *ctx, const double *q, const double *crx, const double *cry, const double *xfx,
               const double *yfx, double *fx, double *fy, const double *x_mass_flux, const double *y_mass_flux,
               const double *mass, int hord, const double *nord_col, const double *damp_col, int nmax, int nk,
               void *stream);
/* ---- DelnFluxNoSG.__call__ (delnflux.py:1209-1261) with mass=None: fx2, fy2 <- del-n fluxes of damp*q */
int fv3_delnflux_nosg(fv3_ctx *ctx, const double *q, double *fx2, double *fy2, const double *damp_col,
                      const double *nord_col, int nmax, int nk, void *stream);

/* ---- AGrid2BGridFourthOrder.__call__ (a2b_ord4.py:673-761) on levels [kstart, kstart+nk); qout != qin */
int fv3_a2b_ord4(fv3_ctx *ctx, const double *qin, double *qout, int kstart, int nk, void *stream);
/* ---- DivergenceDamping.__call__ (divergence_damping.py:482-632); uc/vc are inputs only here (the reference
 *      also uses them as scratch for the Laplacian iterations) */
int fv3_divergence_damping(fv3_ctx *ctx, const double *u, const double *v, const double *va,
                           double *damped_rel_vort_bgrid, const double *ua, double *divg_d, const double *vc,
                           const double *uc, double *delpc, double *ke, const double *rel_vort_agrid, double dt,
                           const fv3_dsw_cols *cols, void *stream);
/* ---- DGridShallowWaterLagrangianDynamics.__call__ (d_sw.py:935-1237), same argument order */
int fv3_d_sw(fv3_ctx *ctx, double *delpc, double *delp, double *pt, double *u, double *v, double *w, double *uc,
             double *vc, const double *ua, const double *va, double *divgd, double *mfx, double *mfy, double *cx,
             double *cy, double *crx, double *cry, double *xfx, double *yfx, double *q_con, const double *zh,
             double *heat_source, double *diss_est, double dt, const fv3_dsw_cols *cols, void *stream);

/* ---- UpdateHeightOnDGrid.__call__ (updatedzd.py:283-356).  gk/beta/gamma: cubic-spline constants
 *      (updatedzd.py:129-154); damp_col: the RAW damp_vt column padded with 0 at level nz (updatedzd.py:337-343);
 *      nord_col: nord_v column; all device double[nz+1]. */
int fv3_update_dz_d(fv3_ctx *ctx, const double *surface_height, double *height, const double *crx, const double *cry,
                    const double *xfx, const double *yfx, double *ws, double dt, const double *gk, const double *beta,
                    const double *gamma, const double *damp_col, const double *nord_col, int nmax, void *stream);

/* ---- NonhydrostaticVerticalSolver.__call__ (riem_solver3.py:207-321), same argument order */
int fv3_riem_solver3(fv3_ctx *ctx, int last_call, double dt, const double *cappa, double ptop, const double *zs,
                     const double *ws, double *delz, const double *q_con, const double *delp, const double *pt,
                     double *zh, double *pe, double *ppe, double *pk3, double *pk, double *peln, double *w,
                     void *stream);
/* ---- edge_pe (pe_halo.py:6-34) and PK3Halo.__call__ (pk3_halo.py:11-69) */
int fv3_edge_pe(fv3_ctx *ctx, double *pe, const double *delp, double ptop, void *stream);
int fv3_pk3_halo(fv3_ctx *ctx, double *pk3, const double *delp, double ptop, double akap, void *stream);
/* ---- NonHydrostaticPressureGradient.__call__ (nh_p_grad.py:190-255); pp, gz, pk3 are replaced by their
 *      B-grid (cell-corner) values as in the reference */
int fv3_nh_p_grad(fv3_ctx *ctx, double *u, double *v, double *pp, double *gz, double *pk3, const double *delp,
                  double dt, double ptop, double akap, void *stream);
/* ---- RayleighDamping.__call__ (ray_fast.py:184-206); rf column and the level counts are host-computed */
int fv3_ray_fast(fv3_ctx *ctx, double *u, double *v, double *w, const double *rf, int n_rf, int n_nudge, double p_ref,
                 void *stream);
/* ---- HyperdiffusionDamping.__call__ (del2cubed.py:165-194) */
int fv3_del2cubed(fv3_ctx *ctx, double *qdel, double cd, int nmax, int nk, void *stream);
/* ---- apply_diffusive_heating (temperature_adjust.py:8-43) on the first nk levels */
int fv3_apply_diffusive_heating(fv3_ctx *ctx, const double *delp, const double *delz, const double *cappa,
                                const double *heat_source, double *pt, double delt_time_factor, int nk, void *stream);

/* ---- LagrangianToEulerian.__call__ (remapping.py:485-695) is issued by the Python class as the calls below.
 *      tracers6: device array of 6 pointers qvapor, qliquid, qrain, qsnow, qice, qgraupel. */
int fv3_remap_prep(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *pt, double *cappa, double *delp,
                   double *delz, const double *pe, double *pe1, double *pe2, double *dp2, double *ps, double *pn2,
                   const double *peln, double *pk, double ptop, double akap, double r_vir, void *stream);
/* MapSingle.__call__ (map_single.py:147-200), kord 9; iv = remap mode (1, 0, -1, -2); qs: NULL, a 2-D field
 * (qs_is_2d = 1) or a 3-D field read at level 0; i_extra / j_extra = 1 for x- / y-interface fields (v / u). */
int fv3_map_single(fv3_ctx *ctx, double *q1, const double *pe1, const double *pe2, const double *qs, int qs_is_2d,
                   double qmin, int kord, int iv, int i_extra, int j_extra, void *stream);
/* n (<= 16) independent MapSingle calls in ONE launch (MapNTracer.__call__, mapn_tracer.py:60-82, and the
 * per-field calls of remapping.py:560-640).  desc: HOST array of n records of 8 int64 {q1, pe1, pe2, qs (0 = none),
 * qs_is_2d, iv, i_extra, j_extra} (device addresses as integers); qmin: HOST array of n doubles.  Both are consumed
 * before the call returns. */
int fv3_map_multi(fv3_ctx *ctx, int n, const int64_t *desc, const double *qmin, int kord, void *stream);
/* FillNegativeTracerValues.__call__ (fillz.py:15-163) for nq tracers (device array of nq pointers) */
int fv3_fillz(fv3_ctx *ctx, double *const *tracers, int nq, const double *dp2, void *stream);
int fv3_remap_post(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *pkz, const double *pt, double *cappa,
                   const double *delp, double *delz, double *peln, double *pe0, const double *pn2, double r_vir,
                   void *stream);
int fv3_remap_pressures(fv3_ctx *ctx, const double *pe, double *pe0, double *pe3, int dir, void *stream);
int fv3_remap_finish(fv3_ctx *ctx, double *const *tracers6, const double *pe2, double *pe, double *pt, const double *pkz,
                     int last_step, double dtmp, double r_vir, void *stream);
/* ---- DynamicalCore.compute_preamble (fv_dynamics.py:440-483): fv_setup + pt_to_potential_density_pt */
int fv3_fv_setup(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *cvm, double *pkz, double *pt,
                 double *cappa, const double *delp, const double *delz, double *dp1, void *stream);
/* ---- AdjustNegativeTracerMixingRatio.__call__ (neg_adj3.py:377-420): fix_neg_water, fillq(qgraupel), fillq(qrain),
 * fix_water_vapor_down, fix_neg_cloud on the compute domain, in place (non-hydrostatic constants) */
int fv3_neg_adj3(fv3_ctx *ctx, double *qvapor, double *qliquid, double *qrain, double *qsnow, double *qice,
                 double *qgraupel, double *qcld, double *pt, const double *delp, void *stream);
/* ---- omega_from_w (fv_dynamics.py:55-64) */
int fv3_omega_from_w(fv3_ctx *ctx, const double *delp, const double *delz, const double *w, double *omga, void *stream);

/* ---- TracerAdvection.__call__ (tracer_2d_1l.py:264-392) pieces; the transport itself is fv3_fvtp2d(hord_tr) */
int fv3_tracer_flux_prep(fv3_ctx *ctx, double *cxd, double *cyd, double *mfxd, double *mfyd, double *xfx, double *yfx,
                         int n_split, void *stream);
int fv3_tracer_apply_mass_flux(fv3_ctx *ctx, const double *dp1, const double *mfx, const double *mfy, double *dp2,
                               void *stream);
int fv3_tracer_apply_flux(fv3_ctx *ctx, double *q, const double *dp1, const double *fx, const double *fy,
                          const double *dp2, void *stream);
int fv3_tracer_swap_dp(fv3_ctx *ctx, double *dp1, double *dp2, void *stream);
/* One whole sub-cycle of the loop at tracer_2d_1l.py:341-392 for nq tracers (device array of pointers): dp2 from the
 * mass fluxes, transport of every tracer with hord (8) and its flux application, then dp2 stored and - when swap != 0,
 * i.e. another sub-cycle follows - exchanged with dp1.  Same results as the four calls above issued per tracer. */
int fv3_tracer_subcycle(fv3_ctx *ctx, double *const *tracers, int nq, double *dp1, double *dp2, const double *mfx,
                        const double *mfy, const double *cx, const double *cy, const double *xfx, const double *yfx,
                        int hord, int swap, void *stream);

/* ---- CubedToLatLon, c2l_ord = 4 (stencils/pace/stencils/c2l_ord.py:41-66); the u/v halo update is done by the caller */
int fv3_c2l_ord4(fv3_ctx *ctx, const double *u, const double *v, double *ua, double *va, void *stream);

/* ---- SatAdjust3d.__call__ (saturation_adjustment.py:945-1108), same argument order (akap is unused there and
 *      dropped); levels [kmp, nz) of the compute domain; te is written only when fast_mp_consv != 0. */
int fv3_sat_adjust(fv3_ctx *ctx, const fv3_sat_adjust_config *config, double *te, double *qvapor, double *qliquid,
                   double *qice, double *qrain, double *qsnow, double *qgraupel, double *qcld, const double *hs,
                   const double *delp, const double *delz, double *q_con, double *pt, double *pkz, double *cappa,
                   double r_vir, double mdt, int fast_mp_consv, int last_step, int kmp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FV3_B200_H */
