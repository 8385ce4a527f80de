#!/usr/bin/env python
"""Generate the committed golden fixtures from the (read-only) reference checkout.

Run ONCE in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [/root/reference]

What it extracts (all data, no code):

* ``mpc_<tree>_<ctrl>.npz`` -- the closed-loop results the reference author saved
  with ``save('savedXXMPCsol.mat')`` (Main.m:107,116,125): first-stage values of the
  871 warm-started QPs (``s_opt .. xi_f_opt``, ``a_opt``), the KPI arrays
  (``P_opt, E_opt, rpm_opt``), the solver flags (``exitMessage``) and the last-step
  dense ``H``/``G`` (RunOpt_ABMPC.m:376-377, RunOpt_FBMPC.m) used to pin the
  QP build + condensing.
* ``nlp_<tree>.npz`` -- the saved NLP trajectory (RunOpt_NLP.m) used to pin the
  RK4 map ``F`` (M=4) and the quintic power map.
* ``lead_TO01_EAD.npz`` -- the lead-vehicle trace ``s_tv, v_tv`` rebuilt with the
  recipe of Run_DrivingCycle.m:13-46 + Main.m:88 (MATLAB ``resample(V,1,5)`` ==
  ``scipy.signal.resample_poly(V,1,5,window=('kaiser',5.0))``, verified to 4.5e-13 m
  through the golden xi_h_opt>0 steps), plus the raw 10 Hz cycle.
* ``lead_WLTC.npz`` -- the same recipe applied to DrivingCycles/WLTC.mat:V_z (highway
  variant of BASELINE config 3).
"""
import os
import sys

import numpy as np
import scipy.io as sio
import scipy.signal as sig

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
TREES = {"mc": "MATLAB_CasADi", "abo": "ACCMPC-ABO_CasADi"}
MPC_FIELDS = ["s_opt", "v_opt", "Fm_opt", "Fb_opt", "xi_v_opt", "xi_h_opt", "xi_s_opt",
              "xi_f_opt", "a_opt", "P_opt", "E_opt", "rpm_opt", "exitMessage", "H", "G"]
NLP_FIELDS = ["s_velInc", "v_velInc", "s_opt", "v_opt", "theta_opt", "j_opt", "Fm_opt",
              "Fb_opt", "P_opt", "E_opt", "a_opt", "rpm_opt",
              "xi_v_opt", "xi_h_opt", "xi_s_opt", "xi_f_opt", "tSolve",
              "cost_P", "cost_a", "cost_j", "cost_xi_v", "cost_xi_h", "cost_xi_s", "cost_xi_f"]


def load_struct(path):
    m = sio.loadmat(path, squeeze_me=True, struct_as_record=False)
    key = [k for k in m if not k.startswith("__")][0]
    return m[key]


def lead_trace(v10hz, t_sim=435.0, ts=0.5, init_dist=10.0, tv_length=4.0):
    """Run_DrivingCycle.m:16-46 and Main.m:88."""
    y = sig.resample_poly(np.asarray(v10hz, dtype=np.float64), 1, 5, window=("kaiser", 5.0))
    y = y.copy()
    y[y < 0.1] = 0.0
    n = int(round(t_sim / ts))
    v = np.zeros(n + 1)
    s = np.zeros(n + 1)
    s[0] = init_dist
    for i in range(1, n):              # MATLAB i = 2..n_cycle
        v[i] = y[i]
        s[i] = s[i - 1] + ts * y[i]
    v[n] = v[n - 1]
    s[n] = s[n - 1]
    return s - tv_length, v


def main():
    for tag, tree in TREES.items():
        for ctrl in ("AB", "FB", "BL"):
            p = os.path.join(REF, tree, f"saved{ctrl}MPCsol.mat")
            if not os.path.exists(p):
                continue
            s = load_struct(p)
            d = {}
            for f in MPC_FIELDS:
                if hasattr(s, f):
                    d[f] = np.asarray(getattr(s, f), dtype=np.float64)
            np.savez_compressed(os.path.join(OUT, f"mpc_{tag}_{ctrl.lower()}.npz"), **d)
            print("wrote", tag, ctrl, {k: v.shape for k, v in d.items() if k in ("H", "G")})
        p = os.path.join(REF, tree, "savedNLPsol.mat")
        s = load_struct(p)
        d = {f: np.asarray(getattr(s, f), dtype=np.float64) for f in NLP_FIELDS}
        np.savez_compressed(os.path.join(OUT, f"nlp_{tag}.npz"), **d)
    cyc = sio.loadmat(os.path.join(REF, TREES["mc"], "DrivingCycles", "TO01_EAD.mat"))
    v_to = cyc["V_TO"].ravel()
    s_tv, v_tv = lead_trace(v_to)
    np.savez_compressed(os.path.join(OUT, "lead_TO01_EAD.npz"), s_tv=s_tv, v_tv=v_tv,
                        v_10hz=v_to.astype(np.float64))
    w = sio.loadmat(os.path.join(REF, TREES["mc"], "DrivingCycles", "WLTC.mat"))
    v_z = w["V_z"].ravel()
    s_w, v_w = lead_trace(v_z, t_sim=435.0)
    np.savez_compressed(os.path.join(OUT, "lead_WLTC.npz"), s_tv=s_w, v_tv=v_w,
                        v_10hz=v_z.astype(np.float64))
    print("lead traces written", s_tv[:3], v_tv[:5])


if __name__ == "__main__":
    main()
