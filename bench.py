#!/usr/bin/env python
"""bench.py -- closed-loop batched MPC throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One bench "step" = one full closed-loop driving cycle (871 MPC steps, each = K5/K4/K2 build ->
K3 QP solve -> post) for this rank's batch of synthetic perturbed scenarios:
    default : BASELINE configs[2], the configuration the metric ("batched FBMPC QP solves/sec") is quoted on:
            FBMPC, 8192 vehicles per GPU (65,536 over the 8 GPUs of a box), perturbed urban + highway scenarios
            with vehicle-mass scatter, full 871-step cycle.
    --ctrl ab --batch 4096 : BASELINE configs[1] "ABMPC closed loop, batch 4096 perturbed urban scenarios, 1 B200"
    N > 1 : weak scaling, the same number of scenarios per GPU, contiguous block partition of the scenario
            index, no data-path collective; one NCCL all_gather of the KPI block at the end of each step.
value = closed-loop QP solves / s with the lead traces and route previews resident in HBM;
e2e   = the same through the C-ABI with HOST buffers (lead upload + KPI download in the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# the sliced pipeline runs up to 16 slices on their own streams: without enough hardware queues (default 8) slices
# share a queue and serialise (measured: 16 slices 612 k QP/s with 8 queues, 706 k with 32).  Must precede CUDA init.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

METRIC = "closed-loop QP solves/sec (batched MPC, full driving cycle)"
UNIT = "QP solves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ctrl", default="fb", choices=["ab", "fb", "bl", "tv"])
    ap.add_argument("--batch", type=int, default=0, help="scenarios per GPU (0 = 8192 for fb, 4096 otherwise)")
    ap.add_argument("--cycle-steps", type=int, default=871, help="MPC steps per driving cycle")
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios per core for the CPU baseline (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nlp", action="store_true", help="skip the K7 (batched NLP benchmark) block of the N = 1 line")
    ap.add_argument("--nlp-seconds", type=float, default=150.0, help="time limit of the K7 block (its stages that finished are kept)")
    ap.add_argument("--parity-scenarios", type=int, default=32, help="scenarios of the timed batch re-checked against the oracle for the `parity` block (0 = skip)")
    ap.add_argument("--groups", type=int, default=0, help="slices the per-step path pipelines the batch as (0 = library default)")
    ap.add_argument("--fused", action="store_true", help="time the fused closed-loop kernel (k5_cycle_kernel) instead of the sliced per-step path")
    ap.add_argument("--solver", default="qpoases", choices=["qpoases", "ocp"],
                    help="closed-loop QP solver: qpoases = dense online active set (K3/K3r, the default and the path the headline is quoted on); "
                         "ocp = stage-structured interior point (K6, the reference's solverToUse = 0 / 2; ab / fb only; ~0.45 s per FBMPC step "
                         "of 8192 vehicles in this first version: use --cycle-steps 40 or so)")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 8192 if args.ctrl == "fb" else 4096
    return args


def lead_2hz():
    d = np.load(os.path.join(GOLDEN, "lead_TO01_EAD.npz"))
    s = d["s_tv"]
    return np.r_[0.0, np.diff(s) / 0.5]


def workload_name(ctrl, B, ns):
    if ctrl == "fb":
        return (f"FBMPC closed loop, {B} vehicles per GPU, perturbed urban (even b) / highway (odd b) lead-vehicle "
                f"scenarios with vehicle-mass scatter, {ns} MPC steps/cycle (BASELINE configs[2]: 65,536 vehicles "
                f"over 8 GPUs = 8192 per GPU)")
    which = "configs[1]" if ctrl == "ab" else "configs[3], full route preview"
    return (f"{ctrl.upper()}MPC closed loop, batch {B} perturbed urban lead-vehicle scenarios per GPU, "
            f"{ns} MPC steps/cycle (BASELINE {which})")


def make_workload(ctrl, B, first, n_steps):
    """Scenarios [first, first+B) of the global synthetic batch (counter-based RNG => rank independent)."""
    from eepacc_mpc_casadi_matlab_b200 import settings as S, usecase as U
    p = S.default_params("mc")
    hw = None
    if ctrl == "fb":
        w = np.load(os.path.join(GOLDEN, "lead_WLTC.npz"))
        hw = np.r_[0.0, np.diff(w["s_tv"]) / 0.5]
    routes, lead = U.perturbed_scenarios(B, lead_2hz(), highway_v_2hz=hw, mass=(ctrl == "fb"),
                                         full_route=(ctrl in ("bl", "tv")), n_steps=n_steps, first=first)
    return p, routes, lead


def synth_inputs(ctrl, B, first):
    """The same batch as make_workload, as inputs of the device-side lead synthesis (k5_lead_synth_kernel)."""
    from eepacc_mpc_casadi_matlab_b200 import usecase as U
    hw = None
    if ctrl == "fb":
        w = np.load(os.path.join(GOLDEN, "lead_WLTC.npz"))
        hw = np.r_[0.0, np.diff(w["s_tv"]) / 0.5]
    pert, _ = U.perturbation_table(B, highway=hw is not None, mass=(ctrl == "fb"),
                                   full_route=(ctrl in ("bl", "tv")), first=first)
    return pert, lead_2hz(), hw


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C port of the reference path; CasADi+qpOASES cannot be built here)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process per core.  Scenario generation and library loading happen BEFORE the barrier; the timer only covers
    the C-side loop over this worker's scenarios (oracle/_build/liboracle_fast.so: -O3 -march=native)."""
    ctrl, first, count, n_steps, barrier = args
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    from oracle import oracle as O
    p, routes, lead = make_workload(ctrl, count, first, n_steps)
    O.run_batch_fast(S.CTRL_NAMES[ctrl], p, routes, lead[:2])        # loads the library, touches the pages
    barrier.wait()
    t0 = time.perf_counter()
    O.run_batch_fast(S.CTRL_NAMES[ctrl], p, routes, lead)
    return t0, time.perf_counter()


def cpu_run(ctrl, n_steps, per_core):
    """All host cores, one process per core, `per_core` scenarios each.  -> (solves/s, cores, sample)."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.build_fast()
    cores = os.cpu_count() or 1
    cores = min(cores, 64)
    ctx = mp.get_context("fork")
    mgr = ctx.Manager()
    barrier = mgr.Barrier(cores)
    jobs = [(ctrl, c * per_core, per_core, n_steps, barrier) for c in range(cores)]
    with ctx.Pool(cores) as pool:
        spans = pool.map(_cpu_worker, jobs, chunksize=1)
    wall = max(t1 for _, t1 in spans) - min(t0 for t0, _ in spans)     # perf_counter is system-wide monotonic
    solves = cores * per_core * n_steps
    return solves / wall, cores, (f"{cores * per_core} scenarios x {n_steps} MPC steps, one process per core, C-side loop "
                                  f"of the -O3 -march=native oracle build, scenario generation outside the clock")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_core = args.cpu_sample or 4
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_run(args.ctrl, min(args.cycle_steps, 60), per_core)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, cores, sample = cpu_run(args.ctrl, args.cycle_steps, per_core)
        vals.append(v)
    wall = time.perf_counter() - t_all
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.ctrl, args.batch, args.cycle_steps),
                   "sample_note": "bounded CPU sample of the same batch (cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/ C port of the reference path (CasADi 3.6.3 + qpOASES 3.2 are binary-only/absent and cannot be built here)",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# parity block of the bench line: the first scenarios of the timed batch, full cycle, CUDA path vs the oracle
# (the oracle only CHECKS here; nothing of it is timed)
# ---------------------------------------------------------------------------------------------
def _parity_worker(args):
    ctrl, first, count, ns = args
    from eepacc_mpc_casadi_matlab_b200 import settings as S
    from oracle import oracle as O
    p, routes, lead = make_workload(ctrl, count, first, ns)
    logs, kpis = [], []
    for b in range(count):
        lg, kp = O.run(S.CTRL_NAMES[ctrl], p, routes[b], lead[:, b, :])
        logs.append(lg); kpis.append(kp)
    return np.stack(logs, 1), np.stack(kpis, 0)


def parity_block(ctrl, n_scen, ns, first, solver="qpoases"):
    import multiprocessing as mp
    from eepacc_mpc_casadi_matlab_b200 import api
    from oracle import oracle as O
    O.build()
    p, routes, lead = make_workload(ctrl, n_scen, first, ns)
    eng = api.Engine(ctrl, p, routes, B=n_scen)
    if solver != "qpoases":
        eng.set_solver(solver)
    if ctrl != "tv":
        eng.set_lead(lead)
    log, kpi = eng.run(ns)
    eng.close()
    cores = min(os.cpu_count() or 1, n_scen, 32)
    per = (n_scen + cores - 1) // cores
    jobs = [(ctrl, first + f, min(per, n_scen - f), ns) for f in range(0, n_scen, per)]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(_parity_worker, jobs)
    ref = np.concatenate([a for a, _ in parts], 1)
    kref = np.concatenate([k for _, k in parts], 0)
    ds = np.abs(log[:, :, :2] - ref[:, :, :2]).max(axis=(0, 2))
    dobj = (np.abs(log[:, :, 11] - ref[:, :, 11]) / (1 + np.abs(ref[:, :, 11]))).max(axis=0)
    e_rel = np.abs(kpi[:, 0] - kref[:, 0]) / (1 + np.abs(kref[:, 0]))
    t_rel = np.abs(kpi[:, 2] - kref[:, 2])
    return {"scenarios_checked": int(n_scen), "steps_per_scenario": int(ns), "checker": "oracle/ (CPU restatement), strict build",
            "frac_within_1e-6": float((ds < 1e-6).mean()), "max_state_err": float(ds.max()),
            "max_force_err_N": float(np.abs(log[:, :, 2:4] - ref[:, :, 2:4]).max()),
            "max_rel_obj_err": float(dobj.max()),
            "status_mismatches": int((log[:, :, 9] != ref[:, :, 9]).sum()),
            "ws_change_count_mismatches": int((log[:, :, 10] != ref[:, :, 10]).sum()),
            "energy_kpi_rel_err_p99": float(np.percentile(e_rel, 99)), "energy_kpi_rel_err_max": float(e_rel.max()),
            "travel_time_kpi_abs_err_max_s": float(t_rel.max()),
            "violation_kpi_mismatches": int((kpi[:, 11:17] != kref[:, 11:17]).sum())}


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def nlp_block(limit_s):
    """K7, the batched NLP benchmark, measured by scripts/bench_nlp.py in a child process AFTER the timed
    region: a stage that does not finish inside the limit is dropped, the headline line is unaffected."""
    import subprocess
    cmd = [sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts", "bench_nlp.py")]
    out = {"what": "k7_nlp_kernel: full-route NLP of RunOpt_NLP.m, one scenario per thread; outside the timed region", "stages": []}
    try:
        pr = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            so, se = pr.communicate(timeout=limit_s)
        except subprocess.TimeoutExpired:
            pr.kill()
            so, se = pr.communicate()
            out["cut_off_after_s"] = limit_s
        for ln in so.splitlines():
            try:
                out["stages"].append(json.loads(ln))
            except ValueError:
                pass
        if pr.returncode not in (0, None, -9) and not out["stages"]:
            out["error"] = (se or "").strip().splitlines()[-1:] or [f"rc {pr.returncode}"]
    except Exception as e:                                            # the block must never cost the headline line
        out["error"] = repr(e)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from eepacc_mpc_casadi_matlab_b200 import api, settings as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, ns = args.batch, args.cycle_steps
    p, routes, lead = make_workload(args.ctrl, B, rank * B, ns)
    eng = api.Engine(args.ctrl, p, routes, B=B, device=local)
    eng.set_groups(args.groups)
    groups = eng.groups()
    n, m = eng.n, eng.m
    lead_pinned = torch.from_numpy(np.ascontiguousarray(lead)).pin_memory()
    eng.set_lead(lead_pinned.numpy())
    from eepacc_mpc_casadi_matlab_b200 import sharding
    kpi_dev = torch.empty((B, S.N_KPI), dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def run_resident():
        l2_flush.zero_()                         # evict whatever the previous cycle left in L2
        _, k = eng.run(ns, want_log=False, device_out=kpi_dev)
        return sharding.gather_kpis(k)            # the only collective: the KPI block of every rank over NCCL/NVLink (no-op at N=1)

    def run_e2e():
        eng.set_lead(lead_pinned.numpy())         # H2D of this step's inputs (pinned host memory)
        _, k = eng.run(ns, want_log=False)         # host KPI buffer: D2H inside the call
        return k

    if args.solver != "qpoases":
        eng.set_solver(args.solver)             # K6 (k6_ocp.cu); the fused kernel and the active-set counters do not apply
    eng.set_fused(args.fused and args.solver == "qpoases")
    for _ in range(max(args.warmup, 0)):
        run_resident()
    # ---- timed: resident inputs ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    cyc_ms = []
    for _ in range(args.steps):
        run_resident()
        if eng.fused() and ns > 1:
            cyc_ms.append(eng.cycle_timing())      # waits for this cycle's fused launch (CUDA events on its stream)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    # ---- instrumented legs (not part of `value`): CUDA events around every kernel of the per-step path.
    # (1) the sliced pipeline: latency of one MPC step of one slice; (2) one cycle as ONE slice, so that every
    # kernel runs alone on the device and its CUDA-event duration is its own ----
    eng.enable_timing(True)
    run_resident()
    step_lat = [eng.timing(ns)["step_ms"].copy()]
    eng.set_groups(1)
    run_resident()
    t = eng.timing(ns)
    kt = {k_: t[k_] for k_ in ("build_ms", "qp_ms", "post_ms")}
    eng.set_groups(groups)
    eng.enable_timing(False)
    ctr = eng.counters()
    # ---- timed: end to end through the C-ABI with host buffers ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        kh = run_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # ---- scenario leg (not part of `value`): the lead traces made in HBM from 128 B per scenario instead of
    # uploaded (k5_lead_synth_kernel); checked against the uploaded traces of the same batch ----
    synth = None
    if rank == 0 and ns >= 2 and args.ctrl != "tv":
        pert, base2, hw2 = synth_inputs(args.ctrl, B, rank * B)
        ts_ = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.synth_lead(pert, base2, hw2, ns)          # returns after the kernel has finished
            ts_.append(time.perf_counter() - t0)
        err = float(np.max(np.abs(eng.fetch_lead() - lead)))
        eng.set_lead(lead_pinned.numpy())
        synth = {"kernel": "k5_lead_synth_kernel", "ms_per_batch": 1e3 * min(ts_), "bytes_written": int(ns * B * 16),
                 "gbs_incl_upload_and_sync": ns * B * 16 / min(ts_) / 1e9, "h2d_bytes": int(pert.nbytes + base2.nbytes),
                 "replaces_h2d_bytes": int(lead_pinned.numel() * 8), "max_abs_diff_vs_host_traces": err}
    # max over ranks
    if world > 1:
        tt = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = tt.tolist()
    else:
        e2e_ms = e2e_s * 1e3
    solves = world * B * ns * args.steps
    value = solves / (ms * 1e-3)
    e2e_val = solves / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst)" if peaks else "fallback 6650 GB/s"
        qp_ms = kt["qp_ms"] / ns
        # algorithmic HBM bytes of one K3 launch (matrix-free closed loop, DESIGN.md "K3"): per QP the
        # structured block (1304 doubles) + g + lba/uba in, hot-start state (x, y, working set) in and
        # out, and x / cost / status / iter out.
        per_qp = (1304 + n + 2 * m) * 8 + 2 * (2 * n * 8 + n * 3 + 8) + (n + 3) * 8
        bytes_qp = B * per_qp
        # FP64 view: F_QP = F_refac + n_iter * F_iter with the DENSE full-space algorithm's flop counts.
        # The reduced-space kernel deliberately executes far fewer (factorisation over ~40 free variables, matrix-free
        # operators): the fraction of the FP64 pipe it really keeps busy is the ncu figure below.
        f_refac = {"fb": 6.0e6, "ab": 3.5e6, "bl": 2.5e5, "tv": 2.5e5}[args.ctrl]
        f_iter = {"fb": 1.55e5, "ab": 1.05e5, "bl": 2.0e4, "tv": 2.0e4}[args.ctrl]
        dfma_peak = api.load_library().eepacc_measure_dfma_peak(local)       # GFLOP/s, measured in this process
        it_mean = float(ctr[:, 0].mean()) / ns          # counters restart with every cycle
        flops_qp = f_refac + it_mean * f_iter
        # DRAM traffic / executed-FP64 share of the dominant kernel from the committed ncu --set full capture
        traffic, exec_frac, cap_note = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k3r_traffic.json")))[args.ctrl]
            traffic = float(tj["dram_bytes_per_vehicle"]) * B if not cyc_ms else None
            exec_frac = tj.get("fp64_pipe_active_frac")
            cap_note = tj.get("capture")
        except Exception:
            pass
        lat = np.concatenate(step_lat)
        tot_k = kt["build_ms"] + kt["qp_ms"] + kt["post_ms"]
        kname = "k3r_qp_kernel (+ k3_qp_kernel for the cold-start step and the hand-overs), per-step path"
        if args.solver == "ocp":
            kname = "k6_ocp_kernel (stage-structured interior point, one vehicle per thread), per-step path"
        launch_ms, share, fused_info = qp_ms, (kt["qp_ms"] / tot_k if tot_k else None), None
        units = B * ns
        if cyc_ms:
            # dominant kernel of the timed region: ONE k5_cycle_kernel launch per cycle (build + QP + post of steps
            # 1..ns-1 of every vehicle)
            kname = "k5_cycle_kernel (fused K2 build + K3r reduced-space QP + K4 post, steps 1..n-1 of every vehicle)"
            launch_ms = float(np.mean([c_["ms"] for c_ in cyc_ms]))
            units = B * (ns - 1)
            share = launch_ms * args.steps / ms if world == 1 else None
            fused_info = {"grid": cyc_ms[-1]["grid"], "ctas_per_sm": cyc_ms[-1]["ctas_per_sm"], "threads_per_cta": 128,
                          "ms_per_launch": launch_ms, "vehicle_steps_per_launch": units,
                          "mean_solve_slot_us": 1e3 * launch_ms * cyc_ms[-1]["grid"] / units,
                          "sm_cycles_per_vehicle_step": {k_: v_ / units for k_, v_ in cyc_ms[-1]["cycles"].items()}}
        # achieved, from the TIMED region: QP solves/s of this GPU x flops per QP (the K3 kernels are >95 % of the
        # GPU time of a step and run concurrently on 16 streams, so a per-launch event time would count co-scheduled
        # launches twice); the single-slice leg gives the per-launch view
        fp64_timed = (value / world) * flops_qp / 1e12                       # TFLOP/s per GPU
        fp64_launch = B * flops_qp / (qp_ms * 1e-3) / 1e12 if not cyc_ms else units * flops_qp / (launch_ms * 1e-3) / 1e12
        hbm_timed = (value / world) * per_qp / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.ctrl, B, ns),
                       "controller": args.ctrl, "batch_per_gpu": B, "cycle_steps": ns, "qp_n": n, "qp_m": m,
                       "path": "fused closed-loop kernel after the cold-start step" if cyc_ms else
                               f"per-step launches, {groups} slices pipelined on their own streams",
                       "solver": args.solver,
                       "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                       "slices": groups,
                       "l2": "a 256 MB buffer is written before every timed cycle (L2 flush); inside a cycle "
                             "every MPC step reads what K2 has just rewritten for that step"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(lead_pinned.numel() * 8),
                    "d2h_bytes_per_step": int(B * S.N_KPI * 8)},
            "gpu_launches": int(launches),
            "lead_synthesis": synth,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": kname, "achieved": fp64_timed, "peak": dfma_peak / 1e3,
                         "unit": "TFLOP/s", "frac": fp64_timed / (dfma_peak / 1e3) if dfma_peak > 0 else None,
                         "traffic": traffic,
                         "achieved_source": "timed region: QP solves/s per GPU x algorithmic flops per QP (SURVEY.md 8(d) "
                                            "dense-algorithm formula F_refac + n_iter F_iter with the measured n_iter)",
                         "flops_per_qp": flops_qp,
                         "peak_source": "measured in this process: register-resident DFMA chains (eepacc_measure_dfma_peak); "
                                        "MEASURED_PEAKS.json holds no FP64 figure",
                         "per_launch": {"achieved": fp64_launch, "frac": fp64_launch / (dfma_peak / 1e3) if dfma_peak > 0 else None,
                                        "ms_per_launch": launch_ms, "share_of_step": share,
                                        "what": "one cycle run as ONE slice, every kernel alone on the device, CUDA events on its stream"},
                         "executed_fp64_pipe_frac": exec_frac, "executed_source": cap_note,
                         "hbm_view": {"achieved": hbm_timed, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_timed / hbm_peak,
                                      "algorithmic_bytes_per_qp": per_qp, "peak_source": peak_src},
                         "note": "warm-started active-set QP, one vehicle per CTA: bound by FP64 / shared-memory LATENCY inside a "
                                 "CTA (dependent chains, barriers), neither by the FP64 pipe's throughput nor by HBM; the dense "
                                 "formula credits ~6e6 flop per QP that the reduced-space kernel does not execute"},
            "fused": fused_info,
            "kernel_ms_per_cycle": dict(kt, what="per-step path, one cycle run as a single slice (kernels serialised)"),
            "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                           "max": float(lat.max()),
                           "what": f"per-step path: one MPC step (build+solve+post) of one slice ({B // groups} vehicles), "
                                   f"CUDA events; {groups} slices run concurrently"},
            "solver": {"mean_ws_changes_per_qp": it_mean, "failed_qps": int(ctr[:, 1].sum())},
        }
        if args.parity_scenarios > 0 and world == 1:
            line["parity"] = parity_block(args.ctrl, min(args.parity_scenarios, B), ns, rank * B, args.solver)
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_run(args.ctrl, ns, args.cpu_sample or 4)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    eng.close()
    if rank == 0:
        if world == 1 and not args.no_nlp and args.impl == "b200":
            line["nlp_benchmark"] = nlp_block(args.nlp_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
