"""`SatAdjust3d` — drop-in for fv3core/pace/fv3core/stencils/saturation_adjustment.py:945-1108: the fast saturation
adjustment (grid-scale condensation / evaporation, freezing, melting, sublimation, cloud fraction) of the vertical
remap, one pointwise CUDA kernel (`fv3_sat_adjust`, pace_b200/csrc/sat_adjust.cu)."""
import ctypes as C

from ... import _lib


class _CConfig(C.Structure):
    _INTS = ["hydrostatic", "rad_snow", "rad_rain", "rad_graupel", "tintqs", "icloud_f"]
    _DBLS = ["sat_adj0", "ql_gen", "qs_mlt", "ql0_max", "t_sub", "qi_gen", "qi_lim", "qi0_max", "dw_ocean", "dw_land", "cld_min",
             "tau_i2s", "tau_v2l", "tau_r2g", "tau_l2r", "tau_l2v", "tau_imlt", "tau_smlt"]
    _fields_ = [(n, C.c_int32) for n in _INTS] + [(n, C.c_double) for n in _DBLS]


class SatAdjust3d:
    def __init__(self, stencil_factory, config, area_64, kmp):
        if config.hydrostatic:
            raise NotImplementedError("Hydrostatic is not implemented")
        self._rt = stencil_factory.runtime
        self._config = config
        self._area_64 = area_64   # read from the runtime's metric-term table on the device
        self._kmp = int(kmp)
        self._c = _CConfig(**{n: int(getattr(config, n)) for n in _CConfig._INTS},
                           **{n: float(getattr(config, n)) for n in _CConfig._DBLS})

    def __call__(self, te, qvapor, qliquid, qice, qrain, qsnow, qgraupel, qcld, hs, peln, delp, delz, q_con, pt, pkz, cappa,
                 r_vir: float, mdt: float, fast_mp_consv: bool, last_step: bool, akap: float, kmp: int):
        """Same arguments as the reference (peln is read by the hydrostatic branch only; akap and kmp are unused
        there too: the levels are those chosen at construction)."""
        self._rt.call("fv3_sat_adjust", C.addressof(self._c), te.ptr, qvapor.ptr, qliquid.ptr, qice.ptr, qrain.ptr, qsnow.ptr,
                      qgraupel.ptr, qcld.ptr, hs.ptr, delp.ptr, delz.ptr, q_con.ptr, pt.ptr, pkz.ptr, cappa.ptr, float(r_vir),
                      float(mdt), int(bool(fast_mp_consv)), int(bool(last_step)), self._kmp)
