"""Multi-GPU plumbing: one process per GPU (torchrun), trajectories sharded by global path id,
per-(start slot, action) log-sum-exp accumulators combined with NCCL all-reduces over NVLink.

Replaces the reference's Distributed.jl worker pool (addprocs / pmap / @spawn,
scripts/distributed_experiments.jl:3,26,28,43-46).  torch is used for exactly two things here:
`torch.distributed` collectives and viewing the library's device buffers as tensors.

A log-sum-exp is not an NCCL reduction op, so the exchange is three collectives around one kernel:
    gmax = allreduce(MAX, lse_max);  lse_sum *= exp(lse_max - gmax) (piql_mc_rescale_dev);
    allreduce(SUM, lse_sum);  allreduce(SUM, count)
"""
import numpy as np


def shard_range(npaths, rank, world):
    """Contiguous block of global path ids [path0, path0 + n) of rank `rank`; blocks tile [0, npaths)."""
    base, rem = divmod(int(npaths), int(world))
    n = base + (1 if rank < rem else 0)
    path0 = rank * base + min(rank, rem)
    return path0, n


class _DevBuf:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def stats_tensors(eng):
    """(lse_max f32, lse_sum f32, count i32) tensor views of the engine's device accumulators."""
    import torch
    m, s, c, n = eng.mc_stats_ptrs()
    dev = "cuda:%d" % eng.device
    return (torch.as_tensor(_DevBuf(m, n, "<f4"), device=dev), torch.as_tensor(_DevBuf(s, n, "<f4"), device=dev),
            torch.as_tensor(_DevBuf(c, n, "<i4"), device=dev))


def allreduce_lse(lse_max, lse_sum, count, rescale, group=None):
    """Combine per-rank (max, sum exp(x - max), count) accumulators in place on every rank.

    rescale(gmax) must bring (lse_max, lse_sum) to the shift gmax: sum *= exp(max - gmax); max = gmax.
    On the GPU that is Engine.mc_rescale (one kernel); the gloo tests pass a torch expression.
    """
    import torch.distributed as dist
    gmax = lse_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    rescale(gmax)
    dist.all_reduce(lse_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return gmax


def allreduce_engine_stats(eng, group=None):
    """The exchange step for one Engine: run on the engine's stream (make it torch's current stream)."""
    m, s, c = stats_tensors(eng)
    allreduce_lse(m, s, c, lambda gmax: eng.mc_rescale(gmax.data_ptr()), group=group)


def mc_logz_sharded(eng, V, beta, npaths, rank, world, seed=0, estimator="A", horizon=1 << 20, starts=None,
                    group=None, want=("sa", "count", "s"), out=None):
    """End-to-end sharded estimate: every rank uploads V, builds the policy table, rolls out its block
    of path ids, joins the all-reduce and downloads the (replicated) result."""
    import torch
    path0, n = shard_range(npaths, rank, world)
    eng.policy_table(V, beta=beta, download=False)
    eng.mc_reset(starts=starts)
    eng.mc_rollout(n, path0=path0, seed=seed, estimator=estimator, horizon=horizon, starts=starts)
    if world > 1:
        with torch.cuda.stream(torch.cuda.ExternalStream(eng.stream, device="cuda:%d" % eng.device)):
            allreduce_engine_stats(eng, group=group)
            torch.cuda.current_stream().synchronize()
    return eng.mc_finalize(want=want, out=out)
