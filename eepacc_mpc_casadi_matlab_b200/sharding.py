"""Multi-GPU layer: scenarios are independent, so the path shards with NO data-path collective.

GPU g of G owns the contiguous block [g*B/G, (g+1)*B/G) of the scenario index;
the only exchange is one all_gather of the per-scenario KPI block (Main.m:191-263 quantities) at
the end of a run -- NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block partition; the first (total % world) ranks get one extra scenario."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_kpis(kpi):
    """kpi: [B_local, N_KPI] tensor on this rank's device -> [sum B_local, N_KPI] on every rank,
    rank-major (== global scenario order under shard_range).  Uneven shards are padded."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return kpi
    world = dist.get_world_size()
    n_local = torch.tensor([kpi.shape[0]], dtype=torch.int64, device=kpi.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = kpi
    if kpi.shape[0] < nmax:
        pad = torch.cat([kpi, kpi.new_zeros((nmax - kpi.shape[0], kpi.shape[1]))])
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def summarize(kpi):
    """Fleet summary of the KPI block: total energy (kWh), mean distance (km), mean travel time at the
    cut-off (s), worst |a| (m/s^2), failed QPs."""
    return {
        "energy_kwh": float(kpi[:, 0].sum() / 3.6e6),
        "mean_distance_km": float(kpi[:, 1].mean() / 1e3),
        "mean_travel_time_s": float(kpi[:, 2].mean()),
        "max_abs_accel": float(torch.maximum(kpi[:, 4].abs(), kpi[:, 5].abs()).max()),
        "failed_qps": int(kpi[:, 7].sum()),
    }
