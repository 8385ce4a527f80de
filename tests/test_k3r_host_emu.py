"""CPU: the LOGIC of the reduced-space QP kernel (csrc/k3r_qp.cu), compiled for the host as a single-thread
emulation (tests/host_emu/k3r_emu.cpp: NT = 1, barriers vanish), replayed on QP operands that K2 produced on a
B200 (tests/golden/qpdump_{ab,fb}_small.npz: structured block, g, bounds and the GPU kernel's answers for two
vehicles over the first 48 closed-loop steps, recorded by scripts/gpu_dump_qps.py).  Test infrastructure only:
the product never loads this library (there is no CPU fallback)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "host_emu"))
import replay as R  # noqa: E402


@pytest.mark.parametrize("cn,fc", [("ab", 48), ("ab", 24), ("fb", 48)])
def test_emulated_kernel_reproduces_recorded_gpu_answers(cn, fc):
    if R.cuda_include() is None:
        pytest.skip("cuda_runtime.h not found (host emulation needs the CUDA headers only)")
    d = np.load(os.path.join(HERE, "golden", f"qpdump_{cn}_small.npz"))
    n, m = int(d["n"]), int(d["m"])
    ns, B = d["x"].shape[:2]
    L = R.build()
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    rp = R.Replayer(L, {"ab": S.AB, "fb": S.FB}[cn], n, m, B, fc)
    deferred = 0
    for k in range(ns):
        dl = rp.solve(d["sblk"][k], d["g"][k], d["lba"][k], d["uba"][k])
        deferred += len(dl)
        assert np.array_equal(rp.status, d["status"][k, :, 0]), k
        ec = np.abs(rp.cost - d["cost"][k, :, 0]) / (1 + np.abs(d["cost"][k, :, 0]))
        assert ec.max() < 1e-9, (k, ec.max())                       # objective: 1e-9 relative
        assert np.abs(rp.x - d["x"][k]).max() < 1e-6, k             # primal: 1e-6 absolute
    if fc < 40:
        assert deferred > 0          # small capacity: the hand-over path (second level / full space) is exercised
