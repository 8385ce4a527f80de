"""GPU: timing probe of the closed loop with the stage-structured solver K6 vs the dense path (short runs).
    python scripts/gpu_ocp_probe.py ctrl B steps"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from eepacc_mpc_casadi_matlab_b200 import api  # noqa: E402
import torch  # noqa: E402

ctrl, B, ns = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
quick = len(sys.argv) > 4 and sys.argv[4] == "quick"      # structured solver only, one slicing
p, routes, lead = bench.make_workload(ctrl, min(B, 8192), 0, ns)
if B > lead.shape[1]:
    rep = B // lead.shape[1]
    lead = np.ascontiguousarray(np.tile(lead, (1, rep, 1)))
    routes = list(routes) * rep
for solver in ("ocp", "qpoases"):
    if solver == "qpoases" and (B > 8192 or quick):
        continue
    eng = api.Engine(ctrl, p, routes, B=B)
    eng.set_solver(solver)
    eng.set_lead(lead)
    for groups in ((1, 4, 16) if solver == "ocp" and not quick else (16,)):
        eng.set_groups(groups)
        eng.run(ns, want_log=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, kpi = eng.run(ns, want_log=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ctr = eng.counters()
        print(f"{ctrl} B={B} steps={ns} solver={solver} groups={eng.groups()}: {dt * 1e3 / ns:.3f} ms per MPC step, "
              f"{B * ns / dt / 1e3:.1f} k QP/s, mean iterations {ctr[:, 0].mean() / ns:.2f}, failed {int(ctr[:, 1].sum())}", flush=True)
    if solver == "ocp":
        eng.set_groups(1)
        eng.enable_timing(True)
        eng.run(ns, want_log=False)
        t = eng.timing(ns)
        print("   kernel ms per step (one slice, CUDA events): build %.3f  qp %.3f  post %.3f" % (t["build_ms"] / ns, t["qp_ms"] / ns, t["post_ms"] / ns))
        eng.enable_timing(False)
    eng.close()
