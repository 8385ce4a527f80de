"""BASELINE config 4: TVMPC / BLMPC with full route preview, batch sweep on one GPU (closed loop, N steps)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for ctrl in ("tv", "bl"):
    for B in [int(b) for b in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1024,4096,16384,65536".split(","))]:
        t0 = time.time()
        p, routes, lead = bench.make_workload(ctrl, B, 0, ns)
        eng = api.Engine(ctrl, p, routes, B=B)
        eng.set_lead(lead)
        kd = torch.empty((B, S.N_KPI), dtype=torch.float64, device="cuda")
        eng.run(ns, want_log=False, device_out=kd)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _, k = eng.run(ns, want_log=False, device_out=kd); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ctr = eng.counters()
        print(json.dumps({"ctrl": ctrl, "batch": B, "cycle_steps": ns, "solves_per_s": B * ns / (ms * 1e-3), "ms_per_mpc_step": ms / ns,
                          "failed_qps": int(ctr[:, 1].sum()), "mean_ws_changes": float(ctr[:, 0].mean() / ns), "gen_s": round(time.time() - t0, 1)}))
        eng.close()
