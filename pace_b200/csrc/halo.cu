// Halo exchange data movement: one launch per exchange for all fields x levels x local subdomains.
// Replaces pack/unpack_scalar_f64, pack/unpack_vector_f64 (util/pace/util/cuda_kernels.py:5-178) and the
// per-field/per-neighbour launch loops of HaloDataTransformerGPU (halo_data_transformer.py:694-915).
// Rotation and vector sign/component swaps are folded into the host-built gather table
// (pace_b200/util/topology.py), so the device side is a pure indexed copy.
#include "common.h"

namespace {
#ifndef FV3_HALO_KB
#define FV3_HALO_KB 8
#endif
constexpr int HALO_KB = FV3_HALO_KB;  // levels moved per thread
}

extern "C" {

int fv3_halo_gather(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *dst_off,
                    const int64_t *src_off, const int8_t *dst_comp, const int8_t *src_comp, const double *sign,
                    int64_t n_entries, void *stream) {
  const int64_t sk = nlev > 1 ? geom->sk : 0;
  // one thread moves HALO_KB levels of one table entry: the entry is decoded once and the level loads are
  // independent requests in flight together
  fv3::launch1d((cudaStream_t)stream, n_entries, (nlev + HALO_KB - 1) / HALO_KB, n_fields, FV_LAMBDA(int64_t e, int kb, int f) {
    const double *src = fields[src_comp[e] * n_fields + f] + src_off[e];
    double *dst = fields[dst_comp[e] * n_fields + f] + dst_off[e];
    const double sg = sign[e];
    const int k0 = kb * HALO_KB;
    double v[HALO_KB];
#pragma unroll
    for (int n = 0; n < HALO_KB; ++n)
      if (k0 + n < nlev) v[n] = src[(k0 + n) * sk];
#pragma unroll
    for (int n = 0; n < HALO_KB; ++n)
      if (k0 + n < nlev) dst[(k0 + n) * sk] = sg * v[n];
  });
  return fv3::check_launch("fv3_halo_gather");
}

int fv3_halo_pack(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *src_off,
                  const int8_t *src_comp, const double *sign, int64_t n_entries, double *buf, void *stream) {
  const int64_t sk = nlev > 1 ? geom->sk : 0;
  fv3::launch1d((cudaStream_t)stream, n_entries, nlev, n_fields, FV_LAMBDA(int64_t e, int k, int f) {
    const double *src = fields[src_comp[e] * n_fields + f];
    buf[((int64_t)f * nlev + k) * n_entries + e] = sign[e] * src[src_off[e] + k * sk];
  });
  return fv3::check_launch("fv3_halo_pack");
}

int fv3_halo_unpack(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *dst_off,
                    const int8_t *dst_comp, int64_t n_entries, const double *buf, void *stream) {
  const int64_t sk = nlev > 1 ? geom->sk : 0;
  fv3::launch1d((cudaStream_t)stream, n_entries, nlev, n_fields, FV_LAMBDA(int64_t e, int k, int f) {
    double *dst = fields[dst_comp[e] * n_fields + f];
    dst[dst_off[e] + k * sk] = buf[((int64_t)f * nlev + k) * n_entries + e];
  });
  return fv3::check_launch("fv3_halo_unpack");
}

// Segmented forms: ONE launch per exchange covers every peer process.  The buffer holds one contiguous segment per peer
// (what NCCL sends / receives); entry e belongs to the segment starting at seg_base[e] (doubles) that holds seg_n[e]
// entries per (field, level), and is the seg_e[e]-th of them:
//   buf[seg_base[e] + (f * nlev + k) * seg_n[e] + seg_e[e]]
int fv3_halo_pack_segments(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *src_off,
                           const int8_t *src_comp, const double *sign, const int64_t *seg_base, const int32_t *seg_n,
                           const int32_t *seg_e, int64_t n_entries, double *buf, void *stream) {
  const int64_t sk = nlev > 1 ? geom->sk : 0;
  fv3::launch1d((cudaStream_t)stream, n_entries, (nlev + HALO_KB - 1) / HALO_KB, n_fields, FV_LAMBDA(int64_t e, int kb, int f) {
    const double *src = fields[src_comp[e] * n_fields + f] + src_off[e];
    const double sg = sign[e];
    const int64_t n = seg_n[e];
    double *dst = buf + seg_base[e] + (int64_t)f * nlev * n + seg_e[e];
    const int k0 = kb * HALO_KB;
    double v[HALO_KB];
#pragma unroll
    for (int q = 0; q < HALO_KB; ++q)
      if (k0 + q < nlev) v[q] = src[(k0 + q) * sk];
#pragma unroll
    for (int q = 0; q < HALO_KB; ++q)
      if (k0 + q < nlev) dst[(k0 + q) * n] = sg * v[q];
  });
  return fv3::check_launch("fv3_halo_pack_segments");
}

int fv3_halo_unpack_segments(const fv3_geom *geom, double *const *fields, int n_fields, int nlev, const int64_t *dst_off,
                             const int8_t *dst_comp, const int64_t *seg_base, const int32_t *seg_n, const int32_t *seg_e,
                             int64_t n_entries, const double *buf, void *stream) {
  const int64_t sk = nlev > 1 ? geom->sk : 0;
  fv3::launch1d((cudaStream_t)stream, n_entries, (nlev + HALO_KB - 1) / HALO_KB, n_fields, FV_LAMBDA(int64_t e, int kb, int f) {
    double *dst = fields[dst_comp[e] * n_fields + f] + dst_off[e];
    const int64_t n = seg_n[e];
    const double *src = buf + seg_base[e] + (int64_t)f * nlev * n + seg_e[e];
    const int k0 = kb * HALO_KB;
    double v[HALO_KB];
#pragma unroll
    for (int q = 0; q < HALO_KB; ++q)
      if (k0 + q < nlev) v[q] = src[(k0 + q) * n];
#pragma unroll
    for (int q = 0; q < HALO_KB; ++q)
      if (k0 + q < nlev) dst[(k0 + q) * sk] = v[q];
  });
  return fv3::check_launch("fv3_halo_unpack_segments");
}

}  // extern "C"
