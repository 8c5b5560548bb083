"""Multi-GPU plumbing: one process per GPU (torchrun), trajectories sharded by global path id,
per-(start slot, action) log-sum-exp accumulators combined with NCCL all-reduces over NVLink.

Replaces the reference's Distributed.jl worker pool (addprocs / pmap / @spawn,
scripts/distributed_experiments.jl:3,26,28,43-46).  torch is used for exactly two things here:
`torch.distributed` collectives and viewing the library's device buffers as tensors.

The exchange step has two implementations:
  * "peer" (default on GPUs): ONE kernel per rank over NVLink peer memory (csrc/peer.cu) -- rank r pulls slice r
    of every rank's three accumulator arrays through CUDA-IPC mappings, combines, and pushes the result into every
    rank's arrays.  torch.distributed only carries the 264-byte handle blobs once per start list.
  * "nccl": a log-sum-exp is not an NCCL reduction op, so it is three collectives around one kernel:
    gmax = allreduce(MAX, lse_max);  lse_sum *= exp(lse_max - gmax) (piql_mc_rescale_dev);
    allreduce(SUM, lse_sum);  allreduce(SUM, count).  Kept as the comparison baseline and for gloo tests.
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


def peer_setup(eng, rank, world, group=None):
    """Exchange the CUDA-IPC handles of the engine's accumulators (after mc_reset) and map the peers'.
    Collective; needed again only when the start list (hence the accumulator allocation) changes."""
    import torch.distributed as dist
    blob = eng.peer_export()
    blobs = [None] * world
    dist.all_gather_object(blobs, blob, group=group)
    eng.peer_import(rank, world, blobs)
    dist.barrier(group=group)   # nobody launches the exchange kernel before every rank has its mappings
    eng._peer_key = (eng.mc_stats_ptrs(), rank, world)


def allreduce_engine_stats(eng, group=None, mode="peer", rank=None, world=None):
    """The exchange step for one Engine, asynchronous on the engine's stream.
    mode "peer": one fused kernel over NVLink peer memory; "nccl": three all-reduces + a rescale kernel (run it
    with the engine's stream as torch's current stream)."""
    if mode == "peer":
        import torch.distributed as dist
        rank = dist.get_rank(group) if rank is None else rank
        world = dist.get_world_size(group) if world is None else world
        if getattr(eng, "_peer_key", None) != (eng.mc_stats_ptrs(), rank, world):
            peer_setup(eng, rank, world, group=group)
        eng.mc_allreduce_peer()
        return
    m, s, c = stats_tensors(eng)
    allreduce_lse(m, s, c, lambda gmax: eng.mc_rescale(gmax.data_ptr()), group=group)


def mc_logz_sharded(eng, V, beta, npaths, rank, world, seed=0, estimator="A", horizon=1 << 20, starts=None,
                    group=None, want=("sa", "count", "s"), out=None, mode="peer", root_only=False):
    """End-to-end sharded estimate: every rank uploads V, builds the policy table, rolls out its block
    of path ids, joins the exchange and downloads the (replicated) result -- with root_only, rank 0 alone
    downloads it (the reference's master collects the results of its workers; the others return stats only)."""
    import torch
    path0, n = shard_range(npaths, rank, world)
    eng.policy_table(V, beta=beta, download=False)
    eng.mc_reset(starts=starts)
    eng.mc_rollout(n, path0=path0, seed=seed, estimator=estimator, horizon=horizon, starts=starts)
    if world > 1:
        with torch.cuda.stream(torch.cuda.ExternalStream(eng.stream, device="cuda:%d" % eng.device)):
            allreduce_engine_stats(eng, group=group, mode=mode, rank=rank, world=world)
            torch.cuda.current_stream().synchronize()
    if root_only and rank != 0:
        return eng.mc_finalize(want=())
    return eng.mc_finalize(want=want, out=out)


class ShardedEstimator:
    """Pipelined form of mc_logz_sharded for callers that iterate (one estimate per V): step(V) enqueues upload, policy
    table, rollout and exchange of this iteration on the engine's stream and hands the accumulators to the library's
    asynchronous finalisation, so the device-to-host copy of iteration k runs under the rollout of iteration k + 1.
    step() returns the result of the PREVIOUS iteration (None at first); finish() returns the last one.  Two sets of host
    buffers alternate; with root_only only rank 0 downloads (the reference's master collects its workers' results)."""

    def __init__(self, eng, rank, world, npaths, mode="peer", group=None, root_only=False, out_dtype=np.float64, buffers=None,
                 estimator="A", horizon=1 << 20):
        self.eng, self.rank, self.world, self.npaths = eng, rank, world, int(npaths)
        self.mode, self.group, self.estimator, self.horizon = mode, group, estimator, horizon
        self.download = not (root_only and rank != 0)
        n = eng.H * eng.W
        cdt = np.uint64 if np.dtype(out_dtype).itemsize == 8 else np.uint32

        def alloc():
            return (np.empty((n, 4), dtype=out_dtype), np.empty((n, 4), dtype=cdt), np.empty((eng.H, eng.W), dtype=out_dtype, order="F"))
        self.bufs = list(buffers) if buffers is not None else [alloc(), alloc()]
        self.k, self.pending = 0, None

    def step(self, V, beta, seed=0):
        import torch
        eng = self.eng
        path0, n = shard_range(self.npaths, self.rank, self.world)
        eng.policy_table_async(V, beta=beta)     # uploaded and built under the previous iteration's rollout
        eng.mc_reset()
        eng.mc_rollout(n, path0=path0, seed=seed, estimator=self.estimator, horizon=self.horizon)
        if self.world > 1:
            with torch.cuda.stream(torch.cuda.ExternalStream(eng.stream, device="cuda:%d" % eng.device)):
                allreduce_engine_stats(eng, group=self.group, mode=self.mode, rank=self.rank, world=self.world)
        prev = self.finish()
        eng.mc_finalize_async(self.bufs[self.k & 1] if self.download else (None, None, None))   # all None: statistics only
        self.pending = self.k
        self.k += 1
        return prev

    def finish(self):
        """Waits for the iteration in flight: ((sa, count, s) or None, stats), or None if nothing is pending."""
        if self.pending is None:
            return None
        out, st = self.eng.mc_finalize_wait()
        self.pending = None
        return (out if self.download else None), st


def solve_sharded(eng, kind, rank, world, V=None, nu=None, beta=1.0, reward_scale=1.0, precision="f64", tol=1e-12,
                  max_sweeps=10 ** 7, check_every=64, x0=None, group=None):
    """Soft value iteration of one grid on `world` GPUs: column slabs, halo columns pushed over NVLink peer memory after every
    round of sweeps, the stop rule's residual summed over the ranks.  Returns (x (H, W) assembled on every rank, sweeps,
    residual); the values are bit-identical to the single-GPU solve for any world size (Jacobi)."""
    import torch
    import torch.distributed as dist
    eng.vi_prepare(kind, V=V, nu=nu, beta=beta, reward_scale=reward_scale, precision=precision)
    if x0 is not None:
        eng.vi_set_iterate(x0)
    blob, (c0, c1) = eng.vi_shard_begin(rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, blob, group=group)
    eng.vi_shard_connect(blobs)
    dist.barrier(group=group)
    sweeps, res = 0, float("inf")
    dev = "cuda:%d" % eng.device if dist.get_backend(group) == "nccl" else "cpu"
    while sweeps < max_sweeps:
        batch = min(check_every, max_sweeps - sweeps)
        eng.vi_sweeps(batch)
        sweeps += batch
        _x, mine = eng.vi_fetch(want_x=False)
        t = torch.tensor([mine], dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=group)
        res = float(t[0])
        if tol > 0 and res < tol:
            break
    x, _ = eng.vi_fetch()
    slabs = [None] * world
    dist.all_gather_object(slabs, (c0, c1, np.ascontiguousarray(x[:, c0:c1])), group=group)
    out = np.full_like(x, np.nan)
    for a, b, blk in slabs:
        out[:, a:b] = blk
    dist.barrier(group=group)     # nobody unmaps while a neighbour may still push
    eng.vi_shard_end()
    return out, sweeps, res


_SHARD_ENV = ("PIQL_TILE_TJ", "PIQL_TILE_NW", "PIQL_TILE_KH", "PIQL_TILE_PLANES_MODE")


def _pin_shape(env):
    import os
    for k in _SHARD_ENV:
        os.environ.pop(k, None)
    for k, v in (env or {}).items():
        os.environ[k] = str(v)


def tune_sharded(eng, kind, rank, world, precision="f32fast", candidates=None, sweeps=32, group=None, prepare=None, keep=True,
                 refine=True, _refining=False):
    """Times the sharded sweeps for a few tile shapes on the devices (every rank must run the same shape, so the library
    cannot tune this by itself: the ranks agree here) and pins the fastest through PIQL_TILE_TJ / _NW / _KH for the solves
    that follow.  candidates: list of dicts of those variables (default: the model's shape and a few shapes sized for a
    rank's share of the tiles, at 4 and 8 sweeps per round).  prepare: callable that (re)does eng.vi_prepare(...) -- needed
    before every piql_vi_shard_begin.  Returns (best env dict, {candidate: us per sweep})."""
    import time
    import torch
    import torch.distributed as dist
    if candidates is None:
        candidates = [{}] + [dict(PIQL_TILE_KH=kh, PIQL_TILE_TJ=tj, PIQL_TILE_NW=nw)
                             for kh in (4, 8) for tj, nw in ((27, 12), (38, 12), (54, 12), (54, 16), (76, 16))]
    prepare = prepare or (lambda: eng.vi_prepare(kind, precision=precision))
    dev = "cuda:%d" % eng.device if dist.get_backend(group) == "nccl" else "cpu"
    results = {}
    for env in candidates:
        _pin_shape(env)
        ok, us = 1.0, float("inf")
        try:
            prepare()
            blob, _cols = eng.vi_shard_begin(rank, world)
        except Exception:
            ok, blob = 0.0, None
        t = torch.tensor([ok], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        if float(t[0]) < 0.5:        # infeasible on some rank (e.g. fewer tile columns than ranks): skipped by everybody
            if blob is not None:
                eng.vi_shard_end()
            continue
        blobs = [None] * world
        dist.all_gather_object(blobs, blob, group=group)
        eng.vi_shard_connect(blobs)
        dist.barrier(group=group)
        if sharded_sweeps_ok(eng, sweeps=8, group=group):
            dist.barrier(group=group)
            t0 = time.perf_counter()
            eng.vi_sweeps(sweeps)
            eng.vi_fetch(want_x=False)
            us = (time.perf_counter() - t0) / sweeps * 1e6
        t = torch.tensor([us], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        results[tuple(sorted(env.items()))] = float(t[0])
        dist.barrier(group=group)
        eng.vi_shard_end()
    best = min(results, key=results.get) if results else ()
    env = dict(best)
    if refine and env.get("PIQL_TILE_TJ") and not _refining:
        # second pass around the winner: neighbouring widths and warp counts at the same depth
        tj, nw, kh = int(env["PIQL_TILE_TJ"]), int(env["PIQL_TILE_NW"]), int(env["PIQL_TILE_KH"])
        near = [dict(PIQL_TILE_KH=kh, PIQL_TILE_TJ=t, PIQL_TILE_NW=w) for t in (tj - 8, tj - 4, tj + 4, tj + 8) for w in (nw - 2, nw, nw + 2)
                if t >= 12 and 4 <= w <= 16]
        env2, res2 = tune_sharded(eng, kind, rank, world, precision=precision, candidates=near, sweeps=sweeps, group=group,
                                  prepare=prepare, keep=False, refine=False, _refining=True)
        results.update(res2)
        best = min(results, key=results.get)
        env = dict(best)
    _pin_shape(env if keep else None)
    return env, results


def sharded_sweeps_ok(eng, sweeps=8, group=None):
    """A few sharded soft-VI sweeps as a canary: True on every rank iff no rank's halo wait timed out (a rank whose kernel
    gave up waiting for a neighbour reports it through vi_fetch).  Every rank makes the same collective call whatever happened
    locally, so the ranks stay in step and can all skip a long timed run together."""
    import torch
    import torch.distributed as dist
    ok = 1.0
    try:
        eng.vi_sweeps(sweeps)
        eng.vi_fetch(want_x=False)
    except Exception:
        ok = 0.0
    dev = "cuda:%d" % eng.device if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([ok], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t[0]) > 0.5
