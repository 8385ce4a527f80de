"""K7 (batched NLP benchmark) on one GPU: one JSON line per stage on stdout, so that a caller with a time
limit (bench.py's `nlp_benchmark` block) keeps whatever finished.  No torch import.  Scenario 0 of every batch is the
reference's own case (urban cycle, MC tree); the others scale the lead's speed by 1 - 3 % steps (no host emulation and no
oracle here: parity is tests/test_gpu_zzz_nlp.py; this script only compares scenario 0 with ipopt's saved optimum)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eepacc_mpc_casadi_matlab_b200 import nlp, settings as S, usecase as U     # noqa: E402

lead = np.load(os.path.join(ROOT, "tests", "golden", "lead_TO01_EAD.npz"))["s_tv"]
gold = np.load(os.path.join(ROOT, "tests", "golden", "nlp_mc.npz"))
p, r = S.default_params("mc"), U.default_route()
tab = nlp.nlp_tables(r, p)


def stage(tag, B, N):
    sc = 1.0 - 0.03 * (np.arange(B) % 4) / 4.0
    s_tv = 6.0 + (lead[None, :N + 1] - 6.0) * sc[:, None]
    t0 = time.time()
    sv = nlp.NlpSolver(p, r, N, B=B, tables=[tab] * B)
    t1 = time.time()
    sol = sv.solve(s_tv)
    t2 = time.time()
    ws = sv.workspace_bytes()
    sv.close()
    out = {"stage": tag, "scenarios": B, "intervals": N, "kernel_ms": sol.ms, "nlp_per_s": B / (sol.ms * 1e-3),
           "interval_iterations_per_s": float(sol.iterations.sum()) * N / (sol.ms * 1e-3),
           "status_counts": {nlp.STATUS[i]: int(c) for i, c in enumerate(np.bincount(sol.status, minlength=5)) if c},
           "iterations": [int(sol.iterations.min()), int(np.median(sol.iterations)), int(sol.iterations.max())],
           "kkt_error_max_of_converged": float(sol.kkt_error[np.isin(sol.status, (0, 4))].max()) if np.isin(sol.status, (0, 4)).any() else None,
           # algorithmic work per interval of one iteration, counted from the code: ~30 kflop, ~4.8 KB
           "roofline_estimate": {"tflops": float(sol.iterations.sum()) * N * 30e3 / (sol.ms * 1e-3) / 1e12,
                                 "gbs": float(sol.iterations.sum()) * N * 4.8e3 / (sol.ms * 1e-3) / 1e9,
                                 "note": "flops / bytes per interval-iteration are code counts, not ncu counters"},
           "workspace_bytes": ws, "setup_s": t1 - t0, "wall_s": t2 - t1,
           "h2d_bytes": int(s_tv.nbytes), "d2h_bytes": int(sol.z.nbytes + sol.f.nbytes + sol.stats.nbytes)}
    if N == 870:
        out["scenario0_vs_ipopt"] = {"max_abs_s_m": float(np.abs(sol.s_opt[0] - gold["s_opt"]).max()),
                                     "max_abs_v_mps": float(np.abs(sol.v_opt[0] - gold["v_opt"]).max()),
                                     "max_abs_Fm_N": float(np.abs(sol.Fm_opt[0] - gold["Fm_opt"]).max()),
                                     "objective": float(sol.f[0]), "ipopt_tSolve_s": float(gold["tSolve"])}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    sms = int(os.environ.get("K7_SMS", 148))
    stage("one warp per SM, 100 intervals", 32 * sms, 100)
    stage("one warp per SM, full cycle (the reference's RunOpt_NLP case)", 32 * sms, 870)
    stage("eight warps per SM, 100 intervals", 8 * 32 * sms, 100)
