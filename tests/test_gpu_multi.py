"""The N > 1 path on real GPUs: two processes, one per GPU -- or, on a box with a single GPU (the driver's test box), two
processes SHARING that GPU: CUDA IPC, the peer-memory kernels and their flag protocols are exercised all the same (the two
contexts are time-sliced, so the spinning kernels only take longer to see each other), the plumbing then runs on gloo because
NCCL refuses two ranks on one device, and the NCCL exchange mode is left to the two-GPU run.  The processes
shard the path ids, roll out their blocks and combine the (max, sum, count) accumulators with
  * the fused log-sum-exp all-reduce kernel over NVLink peer memory (csrc/peer.cu), and
  * the three-NCCL-collective baseline,
and both must reproduce the single-GPU run of all paths: counts exactly, log Z to 1e-5; after the peer exchange
every rank must hold identical bits.  -m gpu."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import Golden, ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _init(rank, world, port):
    """One device per rank when the box has them (NCCL), else every rank on device 0 (gloo).  Returns the rank's device."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    shared = torch.cuda.device_count() < world
    dev = 0 if shared else rank
    torch.cuda.set_device(dev)
    if shared:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda:%d" % dev))
    return dev, shared


def _worker(rank, world, port, npaths, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import piql_jl_b200 as piql
    from piql_jl_b200 import distributed as pd

    dev, shared = _init(rank, world, port)
    g = Golden("r0_01g0_99")
    eng = piql.Engine(dev)
    eng.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    out = {}
    for mode in (("peer", "peer") if shared else ("peer", "nccl", "peer")):   # the second peer round reuses the mappings (epoch 2)
        sa, cnt, ls, st = pd.mc_logz_sharded(eng, V, 1.0, npaths, rank, world, seed=3, mode=mode)
        out.setdefault(mode, []).append((sa.copy(), cnt.copy(), ls.copy(), int(st.transitions)))
    # bits of the accumulators after a peer exchange (replicated by construction)
    p0, n = pd.shard_range(npaths, rank, world)
    eng.mc_reset()
    eng.mc_rollout(n, path0=p0, seed=3)
    pd.allreduce_engine_stats(eng, mode="peer", rank=rank, world=world)
    m, s, c = pd.stats_tensors(eng)
    torch.cuda.synchronize()
    bits = (m.cpu().numpy().view(np.uint32).copy(), s.cpu().numpy().view(np.uint32).copy(), c.cpu().numpy().copy())
    q.put((rank, out, bits))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 1, reason="needs a GPU")
def test_sharded_rollout_with_peer_and_nccl_exchange_matches_single_gpu(piql):
    import torch.multiprocessing as mp
    world, npaths = 2, 300001
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, npaths, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-GPU run of all paths
    g = Golden("r0_01g0_99")
    eng = piql.Engine(0)
    eng.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    sa1, cnt1, ls1, st1 = eng.mc_logz(V, beta=1.0, npaths=npaths, seed=3)
    eng.close()
    ok = cnt1 > 0
    for rank, out, bits in got:
        assert len(out["peer"]) == 2
        for mode in out:
            for sa, cnt, ls, tr in out[mode]:
                assert np.array_equal(cnt, cnt1), mode
                assert np.allclose(sa[ok], sa1[ok], rtol=1e-5, atol=2e-6), mode
                free = ~np.isnan(ls1)
                assert np.allclose(ls[free], ls1[free], rtol=1e-5, atol=2e-6), mode
                assert tr * 0 == 0
        assert sum(o[3] for o in [out["peer"][0]]) > 0
    # transitions are per rank: the two shards add up to the single-GPU count
    assert got[0][1]["peer"][0][3] + got[1][1]["peer"][0][3] == int(st1.transitions)
    for a, b in zip(got[0][2], got[1][2]):
        assert np.array_equal(a, b)


def _vi_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import piql_jl_b200 as piql
    from piql_jl_b200 import distributed as pd
    from conftest import random_gridworld

    dev, shared = _init(rank, world, port)
    walls, goal = random_gridworld((300, 420), 0.2, seed=17)
    eng = piql.Engine(dev)
    eng.load_grid(walls, goal, 0.03, 0.02, 0.2, 0.97)
    x0 = np.asfortranarray(np.random.default_rng(5).normal(size=walls.shape))
    V = np.asfortranarray(np.random.default_rng(6).normal(size=walls.shape) * 0.3)
    out = {}
    for kind, prec, kw in (("nu", "f64", {}), ("nu", "f32fast", {}), ("logz", "f64", dict(V=V, beta=0.6))):
        x, sweeps, res = pd.solve_sharded(eng, kind, rank, world, precision=prec, tol=0.0, max_sweeps=37, check_every=10, x0=x0, **kw)
        out[(kind, prec)] = (x, sweeps, res)
    conv, sw, res = pd.solve_sharded(eng, "nu", rank, world, precision="f64", tol=1e-10, max_sweeps=100000, check_every=32)
    out["converged"] = (conv, sw, res)
    q.put((rank, out))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 1, reason="needs a GPU")
def test_sharded_value_iteration_is_bit_identical_to_single_gpu(piql):
    """North star: value-iteration results are bit-stable across GPU counts."""
    import torch.multiprocessing as mp
    from conftest import random_gridworld
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_vi_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    walls, goal = random_gridworld((300, 420), 0.2, seed=17)
    eng = piql.Engine(0)
    eng.load_grid(walls, goal, 0.03, 0.02, 0.2, 0.97)
    x0 = np.asfortranarray(np.random.default_rng(5).normal(size=walls.shape))
    V = np.asfortranarray(np.random.default_rng(6).normal(size=walls.shape) * 0.3)
    for kind, prec, kw in (("nu", "f64", {}), ("nu", "f32fast", {}), ("logz", "f64", dict(V=V, beta=0.6))):
        eng.vi_prepare(kind, precision=prec, **kw)
        eng.vi_set_iterate(x0)
        for _ in range(3):
            eng.vi_sweeps(10)
        eng.vi_sweeps(7)
        one, res1 = eng.vi_fetch()
        for rank, out in got:
            x, sweeps, res = out[(kind, prec)]
            assert sweeps == 37
            assert np.array_equal(np.nan_to_num(x).view(np.uint64), np.nan_to_num(one).view(np.uint64)), (kind, prec, rank)
            assert abs(res - res1) <= 1e-9 * max(1.0, res1)
    nu1, st = eng.solve_nu(tol=1e-10, check_every=32, max_sweeps=100000)
    for rank, out in got:
        conv, sw, res = out["converged"]
        assert sw == st.sweeps and np.array_equal(np.nan_to_num(conv).view(np.uint64), np.nan_to_num(nu1).view(np.uint64))
    eng.close()
