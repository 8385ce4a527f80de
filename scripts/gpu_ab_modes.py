"""Developer aid (GPU box): closed loop with the reduced-space QP kernel vs the full-space kernel only:
agreement of the logged trajectories and time per MPC step."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload
from eepacc_mpc_casadi_matlab_b200 import api
ctrl, B, ns = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
fcs = [int(a) for a in sys.argv[4:]] or [48]
p, routes, lead = make_workload(ctrl, B, 0, ns)
res = {}
for mode, fc in [(1, 0)] + [(0, f) for f in fcs]:
    eng = api.Engine(ctrl, p, routes, B=B)
    eng.L.eepacc_mem_set_qp_mode(eng.mem, mode, fc)
    eng.set_lead(lead)
    eng.run(min(ns, 30), want_log=False)        # warm-up
    eng.reset()
    eng.enable_timing(True)
    t0 = time.perf_counter()
    log, kpi = eng.run(ns, want_log=True)
    dt = time.perf_counter() - t0
    t = eng.timing(ns)
    ctr = eng.counters()
    res[(mode, fc)] = log
    print(f"{ctrl} B={B} mode={mode} fc={fc}: {B*ns/dt:,.0f} solves/s wall; qp {t['qp_ms']/ns:.3f} ms/step build {t['build_ms']/ns:.3f} post {t['post_ms']/ns:.3f}"
          f" p50 {np.percentile(t['step_ms'],50):.3f} p99 {np.percentile(t['step_ms'],99):.3f} ms; failed {int(ctr[:,1].sum())} mean it {ctr[:,0].mean()/ns:.2f}", flush=True)
    eng.close()
ref = res[(1, 0)]
for k, log in res.items():
    if k == (1, 0):
        continue
    d = np.abs(log[..., :2] - ref[..., :2]).max(axis=(0, 2))          # s, v per vehicle over all steps
    print(f"vs full-space, fc={k[1]}: vehicles with max|d(s,v)| > 1e-6: {(d > 1e-6).sum()} / {B}; > 1e-3: {(d > 1e-3).sum()}; worst {d.max():.3e}; status mismatches {(log[..., 9] != ref[..., 9]).sum()}")
