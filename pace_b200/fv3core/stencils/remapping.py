"""LagrangianToEulerian — drop-in for fv3core/pace/fv3core/stencils/remapping.py:275-695."""
from typing import Dict

import numpy as np
import torch

from ... import constants as c
from ...util.quantity import Quantity

_C3 = (c.X_DIM, c.Y_DIM, c.Z_DIM)
_C3I = (c.X_DIM, c.Y_DIM, c.Z_INTERFACE_DIM)
CONSV_MIN = 0.001
TRACER_ORDER = ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke"]  # gt4py_utils.py:24-34


class LagrangianToEulerian:
    def __init__(self, stencil_factory, quantity_factory, config, area_64, nq: int, pfull, tracers: Dict[str, Quantity],
                 checkpointer=None):
        if config.kord_tm >= 0:
            raise NotImplementedError("map ppm, untested mode where kord_tm >= 0")
        if config.hydrostatic:
            raise NotImplementedError("Hydrostatic is not implemented")
        for k in (abs(config.kord_tm), abs(config.kord_tr), config.kord_wz, config.kord_mt):
            if k != 9:
                raise NotImplementedError("only kord 9 is implemented")
        self._rt = rt = stencil_factory.runtime
        self._checkpointer = checkpointer
        self._do_sat_adjust = bool(config.do_sat_adj)
        # first level whose reference pressure exceeds 10 hPa (remapping.py:344-348)
        nz = rt.comm.geometry.nz
        pf = pfull if pfull is not None else rt.grid_data.p
        pf = getattr(pf, "data", pf)
        pf = pf.detach().cpu().numpy() if isinstance(pf, torch.Tensor) else np.asarray(pf)
        pf = np.asarray(pf, dtype=np.float64).reshape(-1)[:nz]
        self.kmp = nz - 1
        for k in range(nz):
            if pf[k] > 10.0e2:
                self.kmp = k
                break
        if self._do_sat_adjust:
            from .saturation_adjustment import SatAdjust3d

            self._saturation_adjustment = SatAdjust3d(stencil_factory, config.sat_adjust, area_64, self.kmp)
        self._t_min = 184.0
        self._nq = int(nq)
        self._fill = bool(config.fill)
        qf = quantity_factory
        self._pe1, self._pe2, self._pe3, self._pe0 = (qf.zeros(_C3I, "Pa") for _ in range(4))
        self._pe3v, self._pe0v = qf.zeros(_C3I, "Pa"), qf.zeros(_C3I, "Pa")
        self._dp2, self._pn2 = qf.zeros(_C3, "Pa"), qf.zeros(_C3, "Pa")
        self._kord_tm, self._kord_tr = abs(config.kord_tm), abs(config.kord_tr)
        self._kord_wz, self._kord_mt = config.kord_wz, config.kord_mt
        names6 = ["qvapor", "qliquid", "qrain", "qsnow", "qice", "qgraupel"]
        self._t6 = torch.tensor([tracers[n].ptr for n in names6], dtype=torch.int64).to(rt.device)
        self._tq_names = TRACER_ORDER[: self._nq]
        self._tq = torch.tensor([tracers[n].ptr for n in self._tq_names], dtype=torch.int64).to(rt.device)
        self._bound = {n: tracers[n].ptr for n in set(names6) | set(self._tq_names)}

    @staticmethod
    def _desc(q, pe1, pe2, iv, qs=None, qs_2d=True, qmin=0.0, i_extra=0, j_extra=0):
        return ([q.ptr, pe1.ptr, pe2.ptr, qs.ptr if qs is not None else 0, int(qs_2d), int(iv), int(i_extra),
                 int(j_extra)], float(qmin))

    def _map_batch(self, descs, kord):
        """MapSingle over several independent fields in one launch (all kord values of a batch are equal: 9)."""
        d = np.asarray([r for r, _ in descs], dtype=np.int64)
        qmin = np.asarray([m for _, m in descs], dtype=np.float64)
        self._rt.call("fv3_map_multi", len(descs), d.ctypes.data, qmin.ctypes.data, int(kord))

    def __call__(self, tracers: Dict[str, Quantity], pt, delp, delz, peln, u, v, w, cappa, q_con, q_cld, pkz, pk, pe, hs,
                 ps, wsd, ak, bk, dp1, ptop: float, akap: float, zvir: float, last_step: bool, consv_te: float,
                 mdt: float):
        rt = self._rt
        for n, p in self._bound.items():
            if tracers[n].ptr != p:
                raise ValueError("tracers must be the Quantities this object was constructed with")
        t6 = self._t6.data_ptr()
        rt.call("fv3_remap_prep", t6, q_con.ptr, pt.ptr, cappa.ptr, delp.ptr, delz.ptr, pe.ptr, self._pe1.ptr,
                self._pe2.ptr, self._dp2.ptr, ps.ptr, self._pn2.ptr, peln.ptr, pk.ptr, float(ptop), float(akap), float(zvir))
        # map_single(pt), mapn_tracer, map_single(w), map_single(delz) (remapping.py:560-612): independent columns of
        # independent fields, one launch
        batch = [self._desc(pt, peln, self._pn2, 1, qmin=self._t_min)]
        batch += [self._desc(tracers[n], self._pe1, self._pe2, 0) for n in self._tq_names]
        batch += [self._desc(w, self._pe1, self._pe2, -2, qs=wsd), self._desc(delz, self._pe1, self._pe2, 1)]
        self._map_batch(batch, self._kord_tm)
        if self._fill:
            rt.call("fv3_fillz", self._tq.data_ptr(), self._nq, self._dp2.ptr)
        rt.call("fv3_remap_post", t6, q_con.ptr, pkz.ptr, pt.ptr, cappa.ptr, delp.ptr, delz.ptr, peln.ptr, self._pe0.ptr,
                self._pn2.ptr, float(zvir))
        rt.call("fv3_remap_pressures", pe.ptr, self._pe0.ptr, self._pe3.ptr, 0)
        rt.call("fv3_remap_pressures", pe.ptr, self._pe0v.ptr, self._pe3v.ptr, 1)
        self._map_batch([self._desc(u, self._pe0, self._pe3, -1, j_extra=1),
                         self._desc(v, self._pe0v, self._pe3v, -1, i_extra=1)], self._kord_mt)
        dtmp = 0.0
        if last_step:
            if consv_te > CONSV_MIN:
                raise NotImplementedError("We do not support consv_te > 0.001 because that would trigger an allReduce")
            elif consv_te < -CONSV_MIN:
                raise NotImplementedError(f"Unimplemented/untested case consv({consv_te})  < -CONSV_MIN({-CONSV_MIN})")
        if self._do_sat_adjust:
            fast_mp_consv = consv_te > CONSV_MIN
            self._saturation_adjustment(dp1, tracers["qvapor"], tracers["qliquid"], tracers["qice"], tracers["qrain"],
                                        tracers["qsnow"], tracers["qgraupel"], q_cld, hs, peln, delp, delz, q_con, pt, pkz,
                                        cappa, zvir, mdt, fast_mp_consv, last_step, akap, self.kmp)
        rt.call("fv3_remap_finish", t6, self._pe2.ptr, pe.ptr, pt.ptr, pkz.ptr, int(bool(last_step)), dtmp, float(zvir))
