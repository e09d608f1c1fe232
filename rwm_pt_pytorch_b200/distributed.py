"""Multi-GPU plumbing for the sampling path: one process per GPU (`torch.distributed`, NCCL on GPUs, gloo in the CPU
tests).  Chains (RWM) and whole ladders (PT) are independent, so the path shards with NO data-path collective: every
rank runs a contiguous block of units with `chain_id_base` set so that the Philox subsequence of a chain is its GLOBAL
id -- results are identical for 1, 2, 4 or 8 GPUs.  Collectives happen once, after the run: a SUM all-reduce of the fp64
accumulators and an (optional) gather of retained samples."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("accept_count", "chain_steps", "sq_jump_sum", "swap_attempts", "swap_accepts", "sq_beta_jump_sum")


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """(first unit, number of units) of `rank`: contiguous blocks, the remainder spread over the first ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n_units), world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def chain_id_base(first_unit: int, n_temps: int = 1) -> int:
    """Global id of the first chain of a shard (ladders own n_temps consecutive chain ids)."""
    return int(first_unit) * int(n_temps)


def local_statistics(algo) -> Dict[str, float]:
    """fp64 accumulators of one rank's sampler (RandomWalkMH_GPU_Optimized / ParallelTemperingRWM_GPU_Optimized)."""
    b = algo._batch
    post = b.post_burn_in_steps()
    out = dict.fromkeys(STAT_KEYS, 0.0)
    out["accept_count"] = float(b.accept_count.sum().item())
    out["chain_steps"] = float(post * b.n_chains)
    if b.K == 1:
        out["sq_jump_sum"] = float(b.sq_jump_sum.sum().item())
    else:
        out["sq_jump_sum"] = float(b.sq_jump_sum.view(b.L, b.K)[:, 0].sum().item())   # cold chains
        out["swap_attempts"] = float(b.swap_rounds() * (b.K - 1) * b.L)
        acc = b.swap_accepts[:, : b.K - 1].to(torch.float64)
        out["swap_accepts"] = float(acc.sum().item())
        betas = b.beta.view(b.L, b.K)[0].to(torch.float64)
        out["sq_beta_jump_sum"] = float((acc.sum(dim=0) * (betas[:-1] - betas[1:]) ** 2).sum().item())
    return out


def local_statistics_tensor(algo) -> torch.Tensor:
    """The same six accumulators as `local_statistics`, as one fp64 tensor on the sampler's device, computed with device
    ops only (no host synchronisation): what a timed loop enqueues between the kernel and the all-reduce."""
    b = algo._batch
    dev = b.accept_count.device
    post = b.post_burn_in_steps()
    z = torch.zeros((), dtype=torch.float64, device=dev)
    parts = [b.accept_count.sum().to(torch.float64), z + float(post * b.n_chains)]
    if b.K == 1:
        parts += [b.sq_jump_sum.sum(), z, z, z]
    else:
        acc = b.swap_accepts[:, : b.K - 1].to(torch.float64)
        betas = b.beta.view(b.L, b.K)[0].to(torch.float64)
        parts += [b.sq_jump_sum.view(b.L, b.K)[:, 0].sum(), z + float(b.swap_rounds() * (b.K - 1) * b.L), acc.sum(),
                  (acc.sum(dim=0) * (betas[:-1] - betas[1:]) ** 2).sum()]
    return torch.stack([p.to(torch.float64) for p in parts])


def allreduce_statistics_tensor(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a `local_statistics_tensor` (asynchronous on the current stream with NCCL)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def statistics_from_tensor(t: torch.Tensor) -> Dict[str, float]:
    return {k: float(v) for k, v in zip(STAT_KEYS, t.tolist())}


def allreduce_statistics(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """SUM all-reduce of the accumulators over all ranks (the path's only collective besides the sample gather)."""
    t = torch.tensor([float(stats.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: float(v) for k, v in zip(STAT_KEYS, t.tolist())}


def pooled_summary(total: Dict[str, float], n_cold_chain_steps: Optional[float] = None) -> Dict[str, float]:
    """Job-level rates from reduced accumulators."""
    steps = max(total["chain_steps"], 1.0)
    cold_steps = n_cold_chain_steps if n_cold_chain_steps else steps
    out = {"acceptance_rate": total["accept_count"] / steps, "esjd": total["sq_jump_sum"] / max(cold_steps, 1.0)}
    if total["swap_attempts"] > 0:
        out["swap_acceptance_rate"] = total["swap_accepts"] / total["swap_attempts"]
        out["pt_esjd"] = total["sq_beta_jump_sum"] / total["swap_attempts"]
    return out


def gather_samples(local: torch.Tensor, counts, dst: Optional[int] = None, group=None):
    """Gather per-rank sample blocks (units, rows, dim) in global unit order.  `counts[r]` = units of rank r (shards may be
    uneven, so blocks are padded to the largest).  dst=None -> every rank gets the result (all_gather); else only `dst`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if dst is None:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


# ---- public multi-GPU entry points: one process per GPU (torchrun), no data-path collective ---------------------------
def make_sharded(algorithm_cls, total_units: int, *args, group=None, **kwargs):
    """Construct THIS rank's shard of a batched sampler: `total_units` independent chains (RandomWalkMH_GPU_Optimized) or
    ladders (ParallelTemperingRWM_GPU_Optimized) are split into contiguous blocks over the ranks of `group`
    (`shard_range`), the shard runs on the current CUDA device, and its Philox subsequences are the GLOBAL chain ids
    (`chain_id_base` / `ladder_id_base`), so the pooled results are identical for 1, 2, 4 or 8 GPUs.  Without an
    initialised process group the whole batch is one shard.  Returns (sampler, (first_unit, n_units))."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    first, count = shard_range(total_units, rank, world)
    if count < 1:
        raise ValueError(f"rank {rank} of {world} would get no unit of {total_units}")
    is_pt = "ParallelTempering" in getattr(algorithm_cls, "__name__", "")
    kw = dict(kwargs)
    if is_pt:
        kw["num_ladders"], kw["ladder_id_base"] = count, first
    else:
        kw["num_chains"], kw["chain_id_base"] = count, first
    if "device" not in kw and torch.cuda.is_available():
        kw["device"] = torch.device("cuda", torch.cuda.current_device())
    return algorithm_cls(*args, **kw), (first, count)


def finish_sharded(algo, shard, total_units: int, gather: bool = False, dst: Optional[int] = 0, group=None):
    """After `generate_samples` on every rank: SUM all-reduce of the accumulators (-> pooled acceptance / ESJD / swap rates
    of the whole job) and, with gather=True, the gather of the retained samples in global unit order (to `dst`, or to every
    rank with dst=None).  Returns (pooled_summary dict, samples or None)."""
    b = algo._batch
    t = allreduce_statistics_tensor(local_statistics_tensor(algo), group=group)
    total = statistics_from_tensor(t)
    summary = pooled_summary(total, total["chain_steps"] / b.K)
    samples = None
    if gather and b.samples is not None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        per_unit = b.samples.shape[0] // shard[1]                        # stored chains per unit (K with store='all')
        counts = [shard_range(total_units, r, world)[1] * per_unit for r in range(world)]
        samples = gather_samples(b.samples[:, : b.rows_written()], counts, dst=dst, group=group)
    return summary, samples
