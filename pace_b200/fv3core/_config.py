"""`DynamicalCoreConfig` and the nested per-component views the hot path reads.

Field names, defaults and the nested-config properties follow fv3core/pace/fv3core/_config.py:14-476
(namelist defaults from util/pace/util/namelist.py:12-64).  f90nml parsing is out of scope.
"""
import dataclasses
from typing import Tuple


@dataclasses.dataclass(frozen=True)
class SatAdjustConfig:
    """fv3core/pace/fv3core/_config.py:16-40"""
    hydrostatic: bool
    rad_snow: bool
    rad_rain: bool
    rad_graupel: bool
    tintqs: bool
    sat_adj0: float
    ql_gen: float
    qs_mlt: float
    ql0_max: float
    t_sub: float
    qi_gen: float
    qi_lim: float
    qi0_max: float
    dw_ocean: float
    dw_land: float
    icloud_f: int
    cld_min: float
    tau_i2s: float
    tau_v2l: float
    tau_r2g: float
    tau_l2r: float
    tau_l2v: float
    tau_imlt: float
    tau_smlt: float


@dataclasses.dataclass(frozen=True)
class RemappingConfig:
    fill: bool
    kord_tm: int
    kord_tr: int
    kord_wz: int
    kord_mt: int
    do_sat_adj: bool
    sat_adjust: SatAdjustConfig

    @property
    def hydrostatic(self) -> bool:
        return self.sat_adjust.hydrostatic


@dataclasses.dataclass(frozen=True)
class RiemannConfig:
    p_fac: float
    a_imp: float
    use_logp: bool
    beta: float


@dataclasses.dataclass(frozen=True)
class DGridShallowWaterLagrangianDynamicsConfig:
    dddmp: float
    d2_bg: float
    d2_bg_k1: float
    d2_bg_k2: float
    d4_bg: float
    ke_bg: float
    nord: int
    n_sponge: int
    grid_type: int
    d_ext: float
    hord_dp: int
    hord_tm: int
    hord_mt: int
    hord_vt: int
    do_f3d: bool
    do_skeb: bool
    d_con: float
    vtdm4: float
    inline_q: bool
    convert_ke: bool
    do_vort_damp: bool
    hydrostatic: bool


@dataclasses.dataclass(frozen=True)
class AcousticDynamicsConfig:
    tau: float
    k_split: int
    n_split: int
    m_split: int
    delt_max: float
    rf_cutoff: float
    rf_fast: bool
    breed_vortex_inline: bool
    use_old_omega: bool
    riemann: RiemannConfig
    d_grid_shallow_water: DGridShallowWaterLagrangianDynamicsConfig

    nord = property(lambda self: self.d_grid_shallow_water.nord)
    grid_type = property(lambda self: self.d_grid_shallow_water.grid_type)
    hydrostatic = property(lambda self: self.d_grid_shallow_water.hydrostatic)
    hord_tm = property(lambda self: self.d_grid_shallow_water.hord_tm)
    p_fac = property(lambda self: self.riemann.p_fac)
    d_ext = property(lambda self: self.d_grid_shallow_water.d_ext)
    d_con = property(lambda self: self.d_grid_shallow_water.d_con)
    beta = property(lambda self: self.riemann.beta)
    use_logp = property(lambda self: self.riemann.use_logp)


@dataclasses.dataclass
class DynamicalCoreConfig:
    dt_atmos: int = 0
    a_imp: float = 0.0
    beta: float = 0.0
    consv_te: float = 0.0
    d2_bg: float = 0.0
    d2_bg_k1: float = 0.0
    d2_bg_k2: float = 0.0
    d4_bg: float = 0.0
    d_con: float = 0.0
    d_ext: float = 0.0
    dddmp: float = 0.0
    delt_max: float = 0.0
    do_sat_adj: bool = False
    do_vort_damp: bool = False
    fill: bool = False
    hord_dp: int = 0
    hord_mt: int = 0
    hord_tm: int = 0
    hord_tr: int = 0
    hord_vt: int = 0
    hydrostatic: bool = False
    k_split: int = 0
    ke_bg: float = 0.0
    kord_mt: int = 0
    kord_tm: int = 0
    kord_tr: int = 0
    kord_wz: int = 0
    n_split: int = 0
    nord: int = 0
    npx: int = 0
    npy: int = 0
    npz: int = 0
    ntiles: int = 0
    nwat: int = 0
    p_fac: float = 0.0
    rf_cutoff: float = 0.0
    tau: float = 0.0
    vtdm4: float = 0.0
    z_tracer: bool = False
    do_qa: bool = False
    layout: Tuple[int, int] = (1, 1)
    grid_type: int = 0
    do_f3d: bool = False
    inline_q: bool = False
    do_skeb: bool = False
    use_logp: bool = False
    moist_phys: bool = True
    check_negative: bool = False
    c2l_ord: int = 4
    regional: bool = False
    m_split: int = 0
    convert_ke: bool = False
    breed_vortex_inline: bool = False
    use_old_omega: bool = True
    rf_fast: bool = False
    adiabatic: bool = False
    nf_omega: int = 1
    fv_sg_adj: int = -1
    n_sponge: int = 1
    # fast saturation adjustment (NamelistDefaults, util/pace/util/namelist.py:22-51)
    tau_r2g: float = 900.0
    tau_smlt: float = 900.0
    tau_imlt: float = 600.0
    tau_i2s: float = 1000.0
    tau_l2r: float = 900.0
    tau_l2v: float = 300.0
    tau_v2l: float = 90.0
    tau_g2v: float = 900.0
    sat_adj0: float = 0.90
    ql_gen: float = 1.0e-3
    ql_mlt: float = 2.0e-3
    qs_mlt: float = 1.0e-6
    ql0_max: float = 2.0e-3
    t_sub: float = 184.0
    qi_gen: float = 1.82e-6
    qi_lim: float = 1.0
    qi0_max: float = 1.0e-4
    rad_snow: bool = True
    rad_rain: bool = True
    rad_graupel: bool = True
    tintqs: bool = False
    dw_ocean: float = 0.10
    dw_land: float = 0.15
    icloud_f: int = 0
    cld_min: float = 0.05

    @property
    def riemann(self) -> RiemannConfig:
        return RiemannConfig(p_fac=self.p_fac, a_imp=self.a_imp, use_logp=self.use_logp, beta=self.beta)

    @property
    def d_grid_shallow_water(self) -> DGridShallowWaterLagrangianDynamicsConfig:
        return DGridShallowWaterLagrangianDynamicsConfig(
            dddmp=self.dddmp, d2_bg=self.d2_bg, d2_bg_k1=self.d2_bg_k1, d2_bg_k2=self.d2_bg_k2, d4_bg=self.d4_bg,
            ke_bg=self.ke_bg, nord=self.nord, n_sponge=self.n_sponge, grid_type=self.grid_type, d_ext=self.d_ext,
            hord_dp=self.hord_dp, hord_tm=self.hord_tm, hord_mt=self.hord_mt, hord_vt=self.hord_vt,
            do_f3d=self.do_f3d, do_skeb=self.do_skeb, d_con=self.d_con, vtdm4=self.vtdm4, inline_q=self.inline_q,
            convert_ke=self.convert_ke, do_vort_damp=self.do_vort_damp, hydrostatic=self.hydrostatic,
        )

    @property
    def acoustic_dynamics(self) -> AcousticDynamicsConfig:
        return AcousticDynamicsConfig(
            tau=self.tau, k_split=self.k_split, n_split=self.n_split, m_split=self.m_split, delt_max=self.delt_max,
            rf_fast=self.rf_fast, rf_cutoff=self.rf_cutoff, breed_vortex_inline=self.breed_vortex_inline,
            use_old_omega=self.use_old_omega, riemann=self.riemann, d_grid_shallow_water=self.d_grid_shallow_water,
        )

    @property
    def sat_adjust(self) -> SatAdjustConfig:
        return SatAdjustConfig(
            hydrostatic=self.hydrostatic, rad_snow=self.rad_snow, rad_rain=self.rad_rain, rad_graupel=self.rad_graupel,
            tintqs=self.tintqs, sat_adj0=self.sat_adj0, ql_gen=self.ql_gen, qs_mlt=self.qs_mlt, ql0_max=self.ql0_max,
            t_sub=self.t_sub, qi_gen=self.qi_gen, qi_lim=self.qi_lim, qi0_max=self.qi0_max, dw_ocean=self.dw_ocean,
            dw_land=self.dw_land, icloud_f=self.icloud_f, cld_min=self.cld_min, tau_i2s=self.tau_i2s, tau_v2l=self.tau_v2l,
            tau_r2g=self.tau_r2g, tau_l2r=self.tau_l2r, tau_l2v=self.tau_l2v, tau_imlt=self.tau_imlt, tau_smlt=self.tau_smlt)

    @property
    def remapping(self) -> RemappingConfig:
        return RemappingConfig(fill=self.fill, kord_tm=self.kord_tm, kord_tr=self.kord_tr, kord_wz=self.kord_wz,
                               kord_mt=self.kord_mt, do_sat_adj=self.do_sat_adj, sat_adjust=self.sat_adjust)


# dycore_config of driver/examples/configs/baroclinic_c12.yaml:41-88 (do_sat_adj off: SURVEY.md §8d / §8f-1)
BAROCLINIC_C12 = dict(
    ntiles=6, nwat=6, dt_atmos=225, a_imp=1.0, beta=0.0, consv_te=0.0, d2_bg=0.0, d2_bg_k1=0.2, d2_bg_k2=0.1,
    d4_bg=0.15, d_con=1.0, d_ext=0.0, dddmp=0.5, delt_max=0.002, do_sat_adj=False, do_vort_damp=True, fill=True,
    hord_dp=6, hord_mt=6, hord_tm=6, hord_tr=8, hord_vt=6, hydrostatic=False, k_split=1, ke_bg=0.0, kord_mt=9,
    kord_tm=-9, kord_tr=9, kord_wz=9, n_split=1, nord=3, p_fac=0.05, rf_fast=True, rf_cutoff=3000.0, tau=10.0,
    vtdm4=0.06, z_tracer=True, do_qa=True, n_sponge=48,
    # the stock file's saturation-adjustment values (:78-87); used when do_sat_adj=True (the stock setting)
    tau_i2s=1000.0, tau_g2v=1200.0, ql_gen=0.001, ql_mlt=0.002, qs_mlt=0.000001, qi_lim=1.0, dw_ocean=0.1, dw_land=0.15,
    icloud_f=0, tau_l2v=300.0, tau_v2l=90.0, fv_sg_adj=0,
)


def baroclinic_config(nx_tile: int, layout=(1, 1), npz=79, **overrides) -> DynamicalCoreConfig:
    kw = dict(BAROCLINIC_C12)
    kw.update(overrides)
    return DynamicalCoreConfig(npx=nx_tile + 1, npy=nx_tile + 1, npz=npz, layout=tuple(layout), **kw)
