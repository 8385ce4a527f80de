"""K5s probe: synthesise full-cycle lead traces for growing batches (run plain, then under
`ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum -k regex:k5_lead`)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U

base = bench.lead_2hz()
p = S.default_params("mc")
for B in (8192, 65536, 262144):
    rng = np.random.default_rng(B)
    pert = np.zeros((B, S.N_PERT))
    pert[:, 0] = rng.uniform(0.85, 1.15, B); pert[:, 1] = rng.integers(0, 21, B); pert[:, 2] = rng.uniform(6, 30, B)
    for r in range(3):
        pert[:, 3 + 3 * r] = rng.uniform(0, 0.3, B); pert[:, 4 + 3 * r] = rng.uniform(0.01, 0.1, B)
        pert[:, 5 + 3 * r] = rng.uniform(0, 2 * np.pi, B)
    eng = api.Engine("tv", p, U.default_route(), B=B)
    eng.synth_lead(pert, base, None, 871)
    t0 = time.perf_counter()
    eng.synth_lead(pert, base, None, 871)
    dt = time.perf_counter() - t0
    err = float("nan")
    if B <= 65536:
        dev = eng.fetch_lead()
        err = float(np.max(np.abs(dev[:, :64] - U.synth_lead_host(pert[:64], base, None, 871))))
        del dev
    print(f"B={B}: call {1e3 * dt:.3f} ms incl. {pert.nbytes / 1e6:.1f} MB upload + sync, {B * 871 * 16 / 1e6:.0f} MB written, "
          f"max |dev - host| over 64 vehicles {err:.2e}", flush=True)
    eng.close()
