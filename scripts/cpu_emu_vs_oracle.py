"""Developer aid (CPU only): closed loop of the product's device code through the host emulation (tests/host_emu:
K2 build -> K3r reduced-space QP -> K4 post, single-threaded) against the oracle on the bench batch.

    python scripts/cpu_emu_vs_oracle.py [ctrl] [n_scen] [n_steps] [first] [fc]
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "host_emu"))


def worker(args):
    ctrl, b, ns, fc = args
    import bench
    import replay
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    from oracle import oracle as O
    p, routes, lead = bench.make_workload(ctrl, 1, b, ns)
    cid = S.CTRL_NAMES[ctrl]
    a, _ = O.run(cid, p, routes[0], lead[:, 0, :])
    f = replay.emu_run(cid, p, routes[0], ns if ctrl == "tv" else lead[:, 0, :], fc)
    ds = np.abs(a[:, :2] - f[:, :2]).max(axis=1)
    du = np.abs(a[:, 2:4] - f[:, 2:4]).max(axis=1)
    dobj = np.abs(a[:, 11] - f[:, 11]) / (1 + np.abs(a[:, 11]))
    bad = np.nonzero((ds > 1e-7) | (dobj > 1e-8) | (a[:, 9] != f[:, 9]) | (a[:, 10] != f[:, 10]))[0]
    return b, (int(bad[0]) if len(bad) else -1), float(ds.max()), float(du.max()), float(dobj.max()), int((a[:, 10] != f[:, 10]).sum()), int((a[:, 9] != f[:, 9]).sum())


if __name__ == "__main__":
    ctrl = sys.argv[1] if len(sys.argv) > 1 else "fb"
    nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 871
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    fc = int(sys.argv[5]) if len(sys.argv) > 5 else 56
    import replay
    from oracle import oracle as O
    O.build(); replay.build()
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = pool.map(worker, [(ctrl, first + b, ns, fc) for b in range(nsc)], chunksize=1)
    nb = 0
    for b, fs, ds, du, dobj, itd, std in res:
        nb += fs >= 0
        if fs >= 0:
            print(f"scen {b:5d}: first diff step {fs:4d}  max|ds| {ds:.2e}  max|dF| {du:.2e}  max rel dobj {dobj:.2e}  iter diffs {itd:3d}  status diffs {std}")
    print(f"{nb} of {nsc} scenarios differ (emulated product path vs oracle)")
