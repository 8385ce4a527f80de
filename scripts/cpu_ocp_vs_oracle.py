"""Developer aid (CPU only): closed loop with the stage-structured interior-point solver K6 through the host emulation
(tests/host_emu: K2 -> K6 -> K4) against the oracle on the bench batch.

    python scripts/cpu_ocp_vs_oracle.py [ctrl] [n_scen] [n_steps] [first] [option=value ...]
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "host_emu"))


def worker(args):
    ctrl, b, ns, opts = args
    import bench
    import replay
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    from oracle import oracle as O
    p, routes, lead = bench.make_workload(ctrl, 1, b, ns)
    cid = S.CTRL_NAMES[ctrl]
    a, _ = O.run(cid, p, routes[0], lead[:, 0, :])
    f, diag = replay.ocp_emu_run(cid, p, routes[0], lead[:, 0, :], check=False, **opts)
    ds = np.abs(a[:, :2] - f[:, :2]).max()
    dsum = np.abs((a[:, 2] + a[:, 3]) - (f[:, 2] + f[:, 3])).max()
    da = np.abs(a[:, 8] - f[:, 8]).max()
    dobj = (np.abs(a[:, 11] - f[:, 11]) / (1 + np.abs(a[:, 11]))).max()
    return b, float(ds), float(dsum), float(da), float(dobj), float(diag[:, 3].mean()), int(diag[:, 3].max()), int((diag[:, 4] != 0).sum()), int((a[:, 9] != 0).sum())


if __name__ == "__main__":
    ctrl = sys.argv[1] if len(sys.argv) > 1 else "fb"
    nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 871
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    opts = {a.split('=')[0]: float(a.split('=')[1]) for a in sys.argv[5:]}      # solver options, e.g. mu0=1e6
    import replay
    from oracle import oracle as O
    O.build(); replay.build()
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(worker, [(ctrl, first + b, ns, opts) for b in range(nsc)], chunksize=1)
    for r in res:
        flag = "" if (r[1] < 1e-6 and r[4] < 1e-7 and r[7] == 0) else "   <--"
        print("scen %5d: states %.2e  Fm+Fb %.2e N  a %.2e  rel obj %.2e  ipm iters mean %.1f max %d  failed %d (oracle failed %d)%s" % (*r, flag))
    arr = np.array([r[1:] for r in res])
    print("max over %d scenarios: states %.2e  Fm+Fb %.2e  a %.2e  rel obj %.2e; mean iters %.1f; failed QPs %d" % (nsc, arr[:, 0].max(), arr[:, 1].max(), arr[:, 2].max(), arr[:, 3].max(), arr[:, 4].mean(), int(arr[:, 6].sum())))
