"""Host-side mirror of the reference's configuration layer.

* ``Params``  <-> ``OPTsettings`` fields + ``SetVehicleParameters()`` that reach the hot path
  (MATLAB_CasADi/Settings.m:10-252, Functions/Settings/SetVehicleParameters.m:12-58 and the
  ACCMPC-ABO_CasADi copies).  Binary layout == ``eepacc_params`` in include/eepacc.h.
* ``Route``   <-> the PWA route arrays produced by GenerateUseCase.m:50-116
  (``eepacc_route`` in include/eepacc.h).

Nothing here computes on the hot path; it only fills the PODs the C-ABI takes.
"""
import ctypes as C
import math

N_HOR = 20
MAX_SL, MAX_CURV, MAX_SLOPE, MAX_STOP, MAX_TL = 16, 48, 24, 8, 4
N_KPI = 18
AB, FB, BL, TV = 0, 1, 2, 3
CTRL_NAMES = {"ab": AB, "fb": FB, "bl": BL, "tv": TV}

FLAG_DEVICE_PTR = 1
N_PERT = 16                      # eepacc.h EEPACC_N_PERT


class Params(C.Structure):
    _fields_ = [
        ("m", C.c_double), ("L", C.c_double), ("h_g", C.c_double), ("L_f", C.c_double),
        ("c_r", C.c_double), ("R_w", C.c_double), ("phi", C.c_double),
        ("P_m_max", C.c_double), ("T_m_max", C.c_double), ("omega_m_r", C.c_double),
        ("eta_TF", C.c_double), ("lambda_", C.c_double), ("mu", C.c_double), ("g", C.c_double),
        ("zeta_a", C.c_double), ("v_max", C.c_double),
        ("F2", C.c_double), ("p10", C.c_double), ("p01", C.c_double),
        ("W_FB", C.c_double * 7), ("W_AB", C.c_double * 7), ("W_BL", C.c_double * 4),
        ("W_TV", C.c_double * 4), ("W_NLP", C.c_double * 7),
        ("Ts", C.c_double), ("tConstACC_ego", C.c_double), ("tConstACC_tar", C.c_double),
        ("N", C.c_int), ("n_integrate_plant", C.c_int), ("est_ego", C.c_int), ("est_tar", C.c_int),
        ("est_bl", C.c_int), ("est_tv", C.c_int), ("ab_route_rows", C.c_int),
        ("fb_ad_quirk", C.c_int), ("nlp_use_fifth", C.c_int), ("pad0_", C.c_int),
        ("s_goal", C.c_double), ("h_min", C.c_double), ("tau_min", C.c_double),
        ("stopVel", C.c_double), ("stopRefDist", C.c_double), ("stopRefVelSlope", C.c_double),
        ("TLStopRegionSize", C.c_double), ("TLstopVel", C.c_double), ("alpha_TTL", C.c_double),
        ("T_hwp", C.c_double), ("A_hwp", C.c_double), ("G_hwp", C.c_double),
        ("BL_a_low", C.c_double), ("BL_a_high", C.c_double), ("BL_j_low", C.c_double),
        ("BL_j_high", C.c_double),
        ("TV_a_low", C.c_double), ("TV_a_high", C.c_double), ("TV_j_low", C.c_double),
        ("TV_j_high", C.c_double),
        ("b_quadr", C.c_double * 6), ("b_fifth", C.c_double * 21),
        ("s_init", C.c_double), ("v_init", C.c_double), ("a_minus1_init", C.c_double),
        ("cutOffDist", C.c_double),
    ]


class Route(C.Structure):
    _fields_ = [
        ("s_speedLim", C.c_double * MAX_SL), ("v_speedLim", C.c_double * MAX_SL),
        ("s_curv", C.c_double * MAX_CURV), ("curvature", C.c_double * MAX_CURV),
        ("s_slope", C.c_double * MAX_SLOPE), ("slope", C.c_double * MAX_SLOPE),
        ("stopLoc", C.c_double * MAX_STOP),
        ("TLLoc", (C.c_double * 4) * MAX_TL),
        ("mass_scale", C.c_double),
        ("n_sl", C.c_int), ("n_curv", C.c_int), ("n_slope", C.c_int), ("n_stop", C.c_int),
        ("n_tl", C.c_int), ("pad_", C.c_int),
    ]


SPEC_MAX_ZONE, SPEC_MAX_CURVE, SPEC_MAX_SLOPE = 6, 12, 10      # eepacc.h EEPACC_SPEC_MAX_*


class RouteSpec(C.Structure):
    """eepacc_route_spec: what GetUseCase.m / a scenario draw specifies; the device builds the Route from it."""
    _fields_ = [
        ("zones", (C.c_double * 2) * SPEC_MAX_ZONE), ("curves", (C.c_double * 3) * SPEC_MAX_CURVE),
        ("slopes", (C.c_double * 3) * SPEC_MAX_SLOPE), ("stops", C.c_double * MAX_STOP),
        ("tls", (C.c_double * 4) * MAX_TL), ("mass_scale", C.c_double),
        ("n_zone", C.c_int), ("n_curve", C.c_int), ("n_slope", C.c_int), ("n_stop", C.c_int),
        ("n_tl", C.c_int), ("pad_", C.c_int),
    ]


B_QUADR = [185, 1.296e-20, 2.301, 0, 0.003728, -0.000181]                 # Settings.m:227
B_FIFTH = [185, 0.427461350854152, 1.33881428409239, 0,                  # Settings.m:228-236
           0.00357911242725773, -0.000183195720362408, 0,
           2.41017657319644e-07, 1.92524236475911e-07,
           -1.21266753753542e-08, 2.69420315493954e-12,
           -6.73422331977247e-11, -4.65716485427471e-11,
           -3.29148609115376e-11, 8.05602619603684e-12,
           4.07099563902699e-16, 5.70507905296593e-15,
           4.78644673413403e-15, 2.49448524129528e-15,
           1.49184094719691e-15, -5.76405750523528e-16]


def _vehicle(p, tree):
    """SetVehicleParameters.m -- 'mc' = BMW i3, 'abo' = Iveco Daily."""
    if tree == "mc":
        p.m, A_f, c_d, p.L, p.h_g, WD = 1443.0, 2.38, 0.29, 2.57, 0.47, 0.53
        p.c_r, p.R_w = 0.0064, 0.3498
        p.F2 = p.p10 = p.p01 = 0.0
    elif tree == "abo":
        p.m, A_f, c_d, p.L, p.h_g, WD = 2620.0, 4.614, 0.46, 3.52, 1.2, 0.45
        p.c_r, p.R_w = 0.0107, 0.361
        p.F2, p.p10, p.p01 = 1.305072, 0.1154, 0.001764
    else:
        raise ValueError(tree)
    p.L_f = WD * p.L
    p.P_m_max = 125 * 1e3
    p.T_m_max = 250.0
    p.omega_m_r = 4800 / 60 * 2 * math.pi
    p.phi = 9.665 * 1 / p.R_w
    p.v_max = 150 / 3.6
    p.eta_TF = 0.985 * 0.93
    p.lambda_ = 1.05
    p.mu = 0.8
    p.g = 9.81
    p.zeta_a = .5 * c_d * 1.225 * A_f


def default_params(tree="mc"):
    """Settings.m with Main.m:44-52 (useCaseNum=0, Ts=0.5, t_sim=435)."""
    p = Params()
    _vehicle(p, tree)
    w_c, w_f = 0.3, 1e7
    p.W_NLP[:] = [220, w_c * 3e5, w_c * 1e7, 4e5, 8e3, w_f * 9e0, w_f]          # Settings.m:13-28
    p.W_FB[:] = [50, w_c * 3e5, w_c * 1e7, 4e5, 8e3, w_f * 9e0, w_f]            # :31-46
    if tree == "mc":                                                           # MC Settings.m:49-62
        w_c, w_f = 0.1, 1e10
        W = [w_c * 3e5, w_c * 1e7, 8e5, 1e4, w_f * 9e0, w_f]
        p.W_AB[:] = [0.0] + [1e-3 * w for w in W]
        p.ab_route_rows = 1
    else:                                                                      # ABO Settings.m:49-64
        p.W_AB[:] = [1.0] * 7
        p.ab_route_rows = 0
    p.W_BL[:] = [1e2, 0, 0, 1e7]                                               # :65-69
    p.W_TV[:] = [1e2, 0, 0, 1e7]                                               # :72-76
    p.Ts = 0.5
    p.tConstACC_ego, p.tConstACC_tar = 3.0, 5.0                                # :105-106
    p.N = N_HOR
    p.n_integrate_plant = 10                                                   # :109
    p.est_ego = p.est_tar = p.est_bl = p.est_tv = 1                            # :103-104,137,211
    p.fb_ad_quirk = 1
    p.nlp_use_fifth = 1                                                        # :92
    p.s_goal = math.inf                                                        # :141
    p.h_min, p.tau_min = 2.0, 0.5                                              # :201-202
    p.stopVel, p.stopRefDist, p.stopRefVelSlope = 0.2, 100.0, 1.0              # :219-221
    p.TLStopRegionSize, p.TLstopVel, p.alpha_TTL = 2.0, -1.0, 3.34             # :222-224
    p.T_hwp, p.A_hwp = 2.0, 2.0                                                # CreateQP_FB.m:461-463
    p.G_hwp = -0.0246 * 2 + 0.010819
    p.BL_a_low, p.BL_a_high, p.BL_j_low, p.BL_j_high = 3.0, 2.0, 3.0, 1.5      # :129-132
    p.TV_a_low, p.TV_a_high, p.TV_j_low, p.TV_j_high = 1.0, 0.5, 2.0, 0.5      # :213-216
    p.b_quadr[:] = B_QUADR
    p.b_fifth[:] = B_FIFTH
    p.s_init, p.v_init, p.a_minus1_init = 0.0, 0.0, 0.0                        # :83-85
    p.cutOffDist = 3.5e3                                                       # :190
    return p
