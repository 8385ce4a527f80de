"""BASELINE config 4: TVMPC / BLMPC with full route preview, batch sweep over B scenarios and N GPUs, full cycle.

Everything scenario-side happens on the device: the host only draws the vectorised spec table
(usecase.spec_table_fast: 848 B + 128 B per scenario, seconds for 10^6), k5_route_gen_kernel builds the route
blocks and k5_lead_synth_kernel the lead traces in HBM.  Under torchrun every rank takes the contiguous block
[rank * B/N, (rank + 1) * B/N) of the global scenario index (the table is partition independent) and the KPI
blocks are all-gathered over NCCL at the end.   usage: sweep_config4.py [steps] [B1,B2,...] [tv,bl]"""
import json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U, sharding

if __name__ == "__main__":
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 871
    batches = [int(b) for b in (sys.argv[2] if len(sys.argv) > 2 else "1024,4096,16384,65536").split(",")]
    ctrls = (sys.argv[3] if len(sys.argv) > 3 else "tv,bl").split(",")
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    base = bench.lead_2hz()
    p = S.default_params("mc")
    for Bg in batches:
        B = Bg // world
        t0 = time.perf_counter()
        pert, specs = U.spec_table_fast(B, first=rank * B, full_route=True)
        gen_s = time.perf_counter() - t0
        for ctrl in ctrls:
            eng = api.Engine(ctrl, p, None, B=B, device=local)
            t1 = time.perf_counter()
            eng.gen_routes(specs)
            route_ms = 1e3 * (time.perf_counter() - t1)
            t1 = time.perf_counter()
            if ctrl != "tv":
                eng.synth_lead(pert, base, None, ns)
            synth_ms = 1e3 * (time.perf_counter() - t1)
            kd = torch.empty((B, S.N_KPI), dtype=torch.float64, device="cuda")
            timing = B <= 65536
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            eng.enable_timing(timing)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, k = eng.run(ns, want_log=False, device_out=kd)
            allk = sharding.gather_kpis(k)
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms = float(ms.item())
            ctr = eng.counters()
            lat = eng.timing(ns)["step_ms"] if timing else None
            per_veh = ctr[:, 0] / ns
            hist, edges = np.histogram(per_veh, bins=[0, 0.25, 0.5, 1, 2, 4, 8, 16, 1e9])
            if rank == 0:
                print(json.dumps({"ctrl": ctrl, "batch_global": Bg, "n_gpus": world, "batch_per_gpu": B, "cycle_steps": ns,
                                  "solves_per_s": Bg * ns / (ms * 1e-3), "ms_per_cycle": ms,
                                  "step_latency_ms_p50": float(np.percentile(lat, 50)) if timing else None,
                                  "step_latency_ms_p99": float(np.percentile(lat, 99)) if timing else None,
                                  "failed_qps_rank0": int(ctr[:, 1].sum()), "mean_ws_changes_per_qp": float(per_veh.mean()),
                                  "ws_changes_per_qp_hist_by_vehicle": {f"<{e:g}": int(h) for e, h in zip(edges[1:], hist)},
                                  "spec_table_gen_s": round(gen_s, 2), "route_gen_ms_incl_upload": round(route_ms, 2),
                                  "lead_synth_ms": round(synth_ms, 2), "kpi_rows_gathered": int(allk.shape[0]),
                                  "fleet": sharding.summarize(allk)}), flush=True)
            eng.close()
            del kd
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()
