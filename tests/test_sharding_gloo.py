"""CPU, world_size 2 over gloo: the N-rank path (block partition + one KPI all_gather) must give
exactly what one rank computes for the whole batch.  The per-shard compute here is the oracle (the
GPU path itself is covered by the -m gpu tests); what is under test is the host-side sharding logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eepacc_mpc_casadi_matlab_b200 import settings as S, sharding, usecase as U

B_TOTAL, N_STEPS = 5, 12


def _lead_2hz():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "lead_TO01_EAD.npz"))
    return np.r_[0.0, np.diff(d["s_tv"]) / 0.5]


def _kpis(first, count):
    from oracle import oracle as O
    p = S.default_params("mc")
    routes, lead = U.perturbed_scenarios(count, _lead_2hz(), n_steps=25, first=first)
    out = np.zeros((count, S.N_KPI))
    for b in range(count):
        _, out[b] = O.run(S.AB, p, routes[b], lead[:N_STEPS, b, :])
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(B_TOTAL, rank, world)
    local = torch.from_numpy(_kpis(lo, hi - lo))
    allk = sharding.gather_kpis(local)
    if rank == 0:
        q.put((allk.numpy(), sharding.summarize(allk)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_the_batch():
    for total in (1, 5, 8, 4096, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_gather_equals_single_rank(oracle):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got, summary = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    ref = _kpis(0, B_TOTAL)
    assert got.shape == ref.shape and np.array_equal(got, ref)        # bit-identical: no cross-scenario state
    assert summary["failed_qps"] == 0 and summary["energy_kwh"] > 0


# ---- the NLP benchmark (K7) shards the same way; per-shard compute here = the host emulation of the device code ----
def _emu_solve_block(params, routes, s_tv):
    import ctypes as C
    from eepacc_mpc_casadi_matlab_b200 import nlp
    import test_nlp_host_emu as T
    L = C.CDLL(T.EMU_LIB)
    N = s_tv.shape[1] - 1
    zs, fs, sts = [], [], []
    for b in range(s_tv.shape[0]):
        z, st = T.emu_solve(L, params, routes[b], nlp.nlp_tables(routes[b], params), s_tv[b], N)
        zs.append(z); fs.append(st[8]); sts.append(st[:8])
    return nlp.NlpSolution(np.array(zs), np.array(fs), np.array(sts), N, 0.0)


def _nlp_batch():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "lead_TO01_EAD.npz"))["s_tv"][:41]
    return np.stack([6.0 + (d - 6.0) * (1.0 + 0.01 * b) for b in range(5)])


def _nlp_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from eepacc_mpc_casadi_matlab_b200 import nlp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, r = S.default_params("mc"), U.default_route()
    sol, rows = nlp.RunOpt_NLP_sharded(p, r, _nlp_batch(), rank, world, solve=_emu_solve_block)
    if rank == 0:
        q.put((rows, sol.z.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_nlp_two_rank_gather_equals_single_rank():
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import test_nlp_host_emu as T
    from eepacc_mpc_casadi_matlab_b200 import nlp
    if not os.path.exists(T.EMU_LIB):
        os.makedirs(os.path.dirname(T.EMU_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", "/usr/local/cuda/include", "-o", T.EMU_LIB, T.EMU_SRC])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nlp_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    rows, n_local = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, r = S.default_params("mc"), U.default_route()
    ref = nlp.summary_rows(_emu_solve_block(p, [r] * 5, _nlp_batch()))
    assert n_local == 3 and rows.shape == (5, 10) and np.array_equal(rows, ref)      # rank-major = global scenario order
