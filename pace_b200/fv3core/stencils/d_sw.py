"""D-grid shallow-water Lagrangian dynamics — drop-in for fv3core/pace/fv3core/stencils/d_sw.py."""
from typing import Dict

import numpy as np


def get_column_namelist(config, npz: int) -> Dict[str, np.ndarray]:
    """Per-level damping parameters (d_sw.py:611-683), as host arrays of length npz."""
    col = {}
    for name in ("ke_bg", "d_con", "nord"):
        col[name] = np.full(npz, float(getattr(config, name)))
    col["d2_divg"] = np.full(npz, min(0.2, config.d2_bg))
    col["nord_v"] = np.full(npz, min(2.0, col["nord"][0]))
    col["nord_w"] = np.full(npz, col["nord_v"][0])
    col["nord_t"] = np.full(npz, col["nord_v"][0])
    col["damp_vt"] = np.full(npz, float(config.vtdm4) if config.do_vort_damp else 0.0)
    col["damp_w"] = np.full(npz, col["damp_vt"][0])
    col["damp_t"] = np.full(npz, col["damp_vt"][0])

    def set_low_kvals(k):
        for name in ("nord", "nord_w", "d_con"):
            col[name][k] = 0
        col["damp_w"][k] = col["d2_divg"][k]

    def lowest_kvals(k):
        set_low_kvals(k)
        if config.do_vort_damp:
            col["nord_v"][k] = 0
            col["damp_vt"][k] = 0.5 * col["d2_divg"][k]

    if npz == 1 or config.n_sponge < 0:
        col["d2_divg"][0] = config.d2_bg
    else:
        col["d2_divg"][0] = max(0.01, config.d2_bg, config.d2_bg_k1)
        lowest_kvals(0)
        if config.d2_bg_k2 > 0.01:
            col["d2_divg"][1] = max(config.d2_bg, config.d2_bg_k2)
            lowest_kvals(1)
        if config.d2_bg_k2 > 0.05:
            col["d2_divg"][2] = max(config.d2_bg, 0.2 * config.d2_bg_k2)
            set_low_kvals(2)
    return col
