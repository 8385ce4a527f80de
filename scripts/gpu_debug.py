"""Developer aid (GPU box): prints CUDA-vs-oracle differences instead of asserting."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U
from oracle import oracle as O

np.set_printoptions(linewidth=200, precision=6)
lead = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "lead_TO01_EAD.npz"))
L = np.stack([lead["s_tv"], lead["v_tv"]], 1)
p = S.default_params("mc")
r = U.default_route()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for cn in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["ab", "bl", "fb", "tv"]):
    ctrl = S.CTRL_NAMES[cn]
    eng = api.Engine(cn, p, r, B=1)
    V = O.Vehicle(ctrl, p, r)
    # step 0 build comparison
    h, g, a, lba, uba = eng.qpbuild(np.zeros((1, 2)), np.zeros(1), np.array([[L[0, 0], 0.0, 0.0]]), np.zeros(1))
    o = V.step(*L[0])
    Ho, go, Ao, lbo, ubo = O.qp_arrays(V.qp())
    fin = np.isfinite(lbo); fu = np.isfinite(ubo)
    print(cn, "build0: H", np.abs(h[0] - Ho).max(), "g", np.abs(g[0] - go).max(), "A", np.abs(a[0] - Ao).max(),
          "lb", np.abs(lba[0][fin] - lbo[fin]).max(), "ub", np.abs(uba[0][fu] - ubo[fu]).max(),
          "infmatch", np.array_equal(np.isfinite(lba[0]), fin), np.array_equal(np.isfinite(uba[0]), fu))
    t = time.time()
    out = eng.step(L[0:1])
    print(cn, "step0 gpu", out["z1"][0], out["a"], out["status"], out["iter"], "t=%.3f" % (time.time() - t))
    print(cn, "step0 orc", o[:11])
    worst = 0
    for k in range(1, nsteps):
        out = eng.step(L[k:k + 1])
        o = V.step(*L[k])
        d = np.abs(out["z1"][0] - o[:8])
        worst = max(worst, d[:2].max())
        if d[:2].max() > 1e-6 or out["status"][0] != o[9] or k < 4:
            print(cn, k, "dev", d.max(), "gpu", out["z1"][0][:4], out["status"], out["iter"], "orc", o[:4], o[9], o[10])
    print(cn, "worst state dev over", nsteps, "steps:", worst, "counters", eng.counters())
    eng.close()
