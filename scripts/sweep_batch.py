"""BASELINE config 4: TVMPC / BLMPC with full route preview, batch sweep on one GPU (closed loop, N steps).
Scenario tables are drawn by all host cores (counter-based RNG: chunks are independent), the lead traces are
synthesised on the device (k5_lead_synth_kernel).  usage: sweep_batch.py [steps] [B1,B2,...]"""
import ctypes as C
import json, os, sys, time
import multiprocessing as mp
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U


def _chunk(a):
    first, count = a
    pert, routes = U.perturbation_table(count, full_route=True, first=first)
    return pert, bytes(routes)


def tables(B, procs):
    step = max(64, B // (4 * procs))
    jobs = [(f, min(step, B - f)) for f in range(0, B, step)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_chunk, jobs)
    pert = np.concatenate([p for p, _ in parts])
    routes = (S.Route * B)()
    C.memmove(routes, b"".join(r for _, r in parts), C.sizeof(routes))
    return pert, routes


if __name__ == "__main__":
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    batches = [int(b) for b in (sys.argv[2] if len(sys.argv) > 2 else "1024,4096,16384,65536").split(",")]
    procs = min(os.cpu_count() or 1, 32)
    base = bench.lead_2hz()
    p = S.default_params("mc")
    for B in batches:
        t0 = time.time()
        pert, routes = tables(B, procs)           # before any CUDA call of this batch size (fork)
        gen_s = time.time() - t0
        for ctrl in ("tv", "bl"):
            eng = api.Engine(ctrl, p, routes, B=B)
            t1 = time.perf_counter()
            eng.synth_lead(pert, base, None, ns)
            synth_ms = 1e3 * (time.perf_counter() - t1)
            kd = torch.empty((B, S.N_KPI), dtype=torch.float64, device="cuda")
            eng.run(ns, want_log=False, device_out=kd)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _, k = eng.run(ns, want_log=False, device_out=kd); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            ctr = eng.counters()
            print(json.dumps({"ctrl": ctrl, "batch": B, "cycle_steps": ns, "solves_per_s": B * ns / (ms * 1e-3),
                              "ms_per_mpc_step": ms / ns, "failed_qps": int(ctr[:, 1].sum()),
                              "mean_ws_changes": float(ctr[:, 0].mean() / ns), "table_gen_s": round(gen_s, 1),
                              "host_procs": procs, "lead_synth_ms": round(synth_ms, 2)}), flush=True)
            eng.close()
