"""The N > 1 host logic on CPU: world_size-2 gloo processes shard the path ids, reduce their own
records and combine (max, sum, count) with the three-collective log-sum-exp all-reduce that the
GPU ranks run over NCCL.  The oracle plays the part of the rollout + reduction kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import Golden, ROOT


def test_shard_ranges_tile_the_path_ids(piql):
    from piql_jl_b200.distributed import shard_range
    for npaths in (0, 1, 7, 1000, (1 << 27) + 5):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(npaths, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == npaths
            for (p0, n), (p1, _m) in zip(blocks, blocks[1:]):
                assert p0 + n == p1
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1


def _worker(rank, world, port, npaths, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    import piql_jl_b200  # noqa: F401  (host code only; no GPU is touched)
    from piql_jl_b200.distributed import allreduce_lse, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("r0_01g0_99")
    org = g.oracle()
    V = np.nan_to_num(g.V(7, 0))
    table = org.policy_table(V, beta=1.0)
    free = org.free_cells()
    p0, n = shard_range(npaths, rank, world)
    lw, meta = org.rollout(table, V[g.goal], free, n, path0=p0, seed=3)
    mx, sm, cnt, tr = oracle.reduce_paths(lw, meta, len(free), path0=p0)
    m = torch.from_numpy(mx.astype(np.float32).ravel())
    s = torch.from_numpy(sm.astype(np.float32).ravel())
    c = torch.from_numpy(cnt.astype(np.int32).ravel())

    def rescale(gmax):  # what piql_mc_rescale_dev does on the GPU
        s.mul_(torch.where(torch.isinf(m), torch.zeros_like(m), torch.exp(m - gmax)))
        m.copy_(gmax)

    allreduce_lse(m, s, c, rescale)
    t = torch.tensor([tr], dtype=torch.int64)
    dist.all_reduce(t)
    q.put((rank, m.numpy().copy(), s.numpy().copy(), c.numpy().copy(), int(t[0])))
    dist.destroy_process_group()


def test_lse_allreduce_matches_single_shard():
    import torch.multiprocessing as mp
    import oracle
    world, npaths = 2, 60001
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, npaths, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = Golden("r0_01g0_99")
    org = g.oracle()
    V = np.nan_to_num(g.V(7, 0))
    table = org.policy_table(V, beta=1.0)
    free = org.free_cells()
    lw, meta = org.rollout(table, V[g.goal], free, npaths, seed=3)
    mx, sm, cnt, tr = oracle.reduce_paths(lw, meta, len(free))
    lsa_ref, _ = oracle.logz_from_stats(mx, sm, cnt)
    for rank, m, s, c, t in got:
        assert t == tr
        assert np.array_equal(c.reshape(-1, 4), cnt.astype(np.int32))
        with np.errstate(divide="ignore", invalid="ignore"):
            lsa = m.astype(np.float64) + np.log(s.astype(np.float64)) - np.log(c.astype(np.float64))
        ok = cnt.ravel() > 0
        assert np.allclose(lsa[ok], lsa_ref.ravel()[ok], rtol=1e-5, atol=2e-6)
    # every rank holds the same bits after the exchange
    assert np.array_equal(got[0][1], got[1][1]) and np.array_equal(got[0][2], got[1][2])


class _FakeShardEngine:
    """Stands in for piql.Engine in solve_sharded's host logic: every rank sweeps the whole grid with the oracle (so the
    halo exchange is trivially right) but owns, and is asked for, only its slab of columns."""

    def __init__(self, org, rank, world, W):
        self.org, self.rank, self.world, self.W, self.device = org, rank, world, W, 0
        self.calls = []

    def vi_prepare(self, kind, **kw):
        self.kind, self.kw = kind, {k: v for k, v in kw.items() if k in ("V", "nu", "beta", "reward_scale") and v is not None}
        self.x = np.zeros((self.org.H, self.org.W), order="F")
        self.res = float("inf")

    def vi_set_iterate(self, x):
        self.x = np.nan_to_num(np.asfortranarray(x))

    def vi_shard_begin(self, rank, world):
        c0, c1 = self.W * rank // world, self.W * (rank + 1) // world
        self.cols = (c0, c1)
        return b"blob%d" % rank, (c0, c1)

    def vi_shard_connect(self, blobs):
        assert blobs == [b"blob%d" % r for r in range(self.world)]
        self.calls.append("connect")

    def vi_sweeps(self, n):
        for _ in range(n):
            new, _d = self.org.sweep(self.kind, self.x, **self.kw)
            c0, c1 = self.cols
            mine = np.nan_to_num(new[:, c0:c1]) - np.nan_to_num(self.x[:, c0:c1])
            self.res = float(np.abs(mine).sum())       # this rank's share of the stop rule
            self.x = np.nan_to_num(new)

    def vi_fetch(self, want_x=True):
        out = self.x.copy()
        out[self.org.walls != 0] = np.nan
        return (out if want_x else None), self.res

    def vi_shard_end(self):
        self.calls.append("end")


def _shard_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import piql_jl_b200  # noqa: F401
    from piql_jl_b200.distributed import solve_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("r0_05g0_95")
    eng = _FakeShardEngine(g.oracle(), rank, world, g.W)
    x, sweeps, res = solve_sharded(eng, "nu", rank, world, tol=1e-9, max_sweeps=4000, check_every=16)
    q.put((rank, x, sweeps, res, eng.calls))
    dist.destroy_process_group()


def test_sharded_solve_host_logic_stitches_slabs_and_sums_the_residual():
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = Golden("r0_05g0_95")
    for rank, x, sweeps, res, calls in got:
        assert calls == ["connect", "end"] and sweeps % 16 == 0 and res < 1e-9
        assert np.array_equal(np.isnan(x), g.walls != 0)
        assert float(np.nanmax(np.abs(x - g.nu))) < 1e-8          # the reference's stored nu
    assert got[0][2] == got[1][2] and np.array_equal(np.nan_to_num(got[0][1]), np.nan_to_num(got[1][1]))


def _canary_worker(rank, world, port, bad_rank, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import piql_jl_b200  # noqa: F401
    from piql_jl_b200.distributed import sharded_sweeps_ok

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Eng:
        device = 0
        swept = 0

        def vi_sweeps(self, n):
            self.swept += n

        def vi_fetch(self, want_x=True):
            if rank == bad_rank:
                raise RuntimeError("a soft-VI kernel timed out waiting for a neighbouring tile or rank")
            return None, 0.0
    eng = Eng()
    q.put((rank, sharded_sweeps_ok(eng, sweeps=8), eng.swept))
    dist.destroy_process_group()


@pytest.mark.parametrize("bad_rank", [-1, 1])
def test_sharded_canary_agrees_on_every_rank(bad_rank):
    """bench.py runs a few sharded sweeps before the long timed run: if any rank's halo wait timed out, all ranks skip it."""
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_canary_worker, args=(r, world, port, bad_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [bad_rank < 0] * world and all(g[2] == 8 for g in got)
