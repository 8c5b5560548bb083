"""Bit-exact oracle parity for the kernel instantiations that produce the headline numbers (BASELINE configs 4 and 5):

  * `rollout_kernel<0, INJ=false, TRACE=false, STAGED=false>` on the 1026^2 and 4098^2 policy tables (34 MB / 537 MB, global
    memory, lanes refilled many times per launch) -- per-path records (log-weight bits, first action, length) against
    `oracle.rollout` on the table the GPU built, for the first 2^20 path ids, for a block at path0 = 3 * 2^27 + 5, and for
    blocks cut out of the middle and the end of a 2^24-path launch (the warp's path ranges and the refill order at scale);
  * the per-(start cell, action) accumulators of that launch against `oracle.reduce_paths` for a sample of start slots;
  * `tile_sweep_kernel` on the 1026^2 / 4098^2 grids: N tiled sweeps from a random iterate against N Jacobi sweeps of the
    oracle (1e-12 in Float64, 2e-6 in fp32), N not a multiple of the temporal-blocking depth.

The smaller grids of tests/test_gpu_parity.py stage the table in shared memory (another instantiation).  -m gpu."""
import os
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

SEED = 20261019


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


class World:
    def __init__(self, engine, name, sweeps):
        import bench
        self.walls, self.goal, self.params = bench.make_world(name)
        engine.load_grid(self.walls, self.goal, **self.params)
        self.org = oracle.Grid(self.walls, self.goal, **self.params)
        self.free = self.org.free_cells()
        V, _ = engine.solve_nu(reward_scale=0.1, tol=0.0, max_sweeps=sweeps, precision="f64")   # the bench's V_lo recipe
        self.V = np.nan_to_num(V)
        self.table = engine.policy_table(self.V, beta=1.0)
        # grids of 2^22 cells and more: the production kernel reads the 16-byte form of the same records
        self.table16, self.rec16 = engine.policy_table16()
        self.vgoal = float(self.V[self.goal])
        self.threads = oracle.num_threads()

    def oracle_records(self, path0, n):
        table = self.table16 if self.rec16 else self.table
        return self.org.rollout(table, self.vgoal, self.free, n, path0=path0, seed=SEED, nthreads=self.threads)


@pytest.fixture(scope="module", params=["grid1024", "grid4096"])
def world(request, engine):
    return World(engine, request.param, 200 if request.param == "grid1024" else 64)


def test_policy_table_bits(engine, world):
    """The 32-byte records are the oracle's, bit for bit, at 1026^2 and 4098^2; the 16-byte records the production kernel reads
    at 4098^2 decode to exactly their running sums, blocked-move bits and first fitness entry."""
    ref = world.org.policy_table(world.V, beta=1.0, nthreads=world.threads)
    assert np.array_equal(world.table.view(np.uint32), ref.view(np.uint32))
    assert world.rec16 == (world.walls.size >= 1 << 22)
    if world.rec16:
        t32, t16 = world.table.view(np.uint32), world.table16
        free = world.free
        assert np.array_equal(t16[free, 0] & 0x3fffffff, t32[free, 0]) and np.array_equal(t16[free, 1] & 0x3fffffff, t32[free, 1])
        assert np.array_equal(t16[free, 2], t32[free, 2]) and np.array_equal(t16[free, 3], t32[free, 4])
        assert np.array_equal((t16[free, 0] >> 30) | ((t16[free, 1] >> 30) << 2), t32[free, 3])
        # with the V-determined Q the four fitness entries agree to the last bits of the double they were rounded from
        bf = world.table[free, 4:8]
        assert float(np.max(np.abs(bf - bf[:, :1]))) <= 1e-6 * max(1.0, float(np.max(np.abs(bf))))


def test_record_forms_sample_the_same_trajectories(engine, world):
    """16- and 32-byte records carry the same running sums and blocked-move bits: first actions and path lengths are
    identical, the log-weights agree to float rounding (the one fitness float against four that agree to ~1 ulp)."""
    if not world.rec16:
        pytest.skip("the 32-byte records are the only form at this size")
    n = 1 << 18
    a = world.org.rollout(world.table16, world.vgoal, world.free, n, path0=11, seed=SEED, nthreads=world.threads)
    b = world.org.rollout(world.table, world.vgoal, world.free, n, path0=11, seed=SEED, nthreads=world.threads)
    assert np.array_equal(a[1], b[1])
    assert np.allclose(a[0], b[0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("path0", [0, 3 * (1 << 27) + 5, (1 << 33) + 12345])
def test_production_rollout_records_bit_exact(engine, world, path0):
    n = 1 << 20
    engine.mc_reset()
    engine.mc_rollout(n, path0=path0, seed=SEED)
    logw, meta = engine.mc_records(n)
    lw, me = world.oracle_records(path0, n)
    assert np.array_equal(meta, me)
    assert np.array_equal(_bits(logw), _bits(lw))
    # and the accumulators of that launch, for a sample of start slots
    sa, cnt, ls, st = engine.mc_finalize()
    mx, sm, cn, tr = oracle.reduce_paths(lw, me, len(world.free), path0=path0)
    assert int(st.transitions) == tr
    lsa, lss = oracle.logz_from_stats(mx, sm, cn)
    cells = world.free
    assert np.array_equal(cnt[cells], cn)
    got = sa[cells]
    ok = cn > 0
    assert np.allclose(got[ok], lsa[ok], rtol=1e-5, atol=2e-6)
    assert np.all(np.isnan(got[~ok]))


def test_production_rollout_inside_a_large_launch(engine, world):
    """One launch of 2^24 paths (14 k paths per warp at 1184 CTAs: every lane refilled ~440 times): blocks from the start,
    the middle and the end of the launch are the oracle's paths; a warp boundary falls inside each block."""
    n, blk = 1 << 24, 1 << 17
    path0 = 7 * (1 << 27)
    engine.mc_reset()
    engine.mc_rollout(n, path0=path0, seed=SEED)
    logw, meta = engine.mc_records(n)
    for off in (0, n // 2 - 777, n - blk):
        lw, me = world.oracle_records(path0 + off, blk)
        assert np.array_equal(meta[off:off + blk], me), off
        assert np.array_equal(_bits(logw[off:off + blk]), _bits(lw)), off


def _sweeps(org, kind, x, n, **kw):
    for _ in range(n):
        x, d = org.sweep(kind, np.nan_to_num(x), **kw)
    return x, d


@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 2e-6), ("f32fast", 4e-6)])
def test_tiled_sweeps_equal_oracle_jacobi_at_full_size(engine, world, prec, tol):
    big = world.walls.shape[0] > 2000
    rng = np.random.default_rng(3)
    x0 = np.asfortranarray(rng.normal(size=world.walls.shape) * 0.5)
    V = np.asfortranarray(rng.normal(size=world.walls.shape) * 0.3)
    nsw = 5 if big else 8 + 1          # crosses a round boundary of the temporal blocking
    cases = [("nu", {})] if big else [("nu", {}), ("logz", dict(V=V, beta=0.4)), ("gap", dict(V=V, nu=world.V))]
    free = world.walls == 0
    cache = world.__dict__.setdefault("sweep_refs", {})   # the oracle's sweeps are shared by the three precisions
    for kind, kw in cases:
        if kind not in cache:
            cache[kind] = _sweeps(world.org, kind, x0, nsw, **kw)
        ref, dref = cache[kind]
        engine.vi_prepare(kind, precision=prec, **kw)
        engine.vi_set_iterate(x0)
        engine.vi_sweeps(nsw)
        out, d = engine.vi_fetch()
        assert np.array_equal(np.isnan(out), ~free)
        err = float(np.max(np.abs(out[free] - ref[free]) / np.maximum(1.0, np.abs(ref[free]))))
        assert err < tol, (kind, prec, err)
        if prec == "f64":
            assert abs(d - dref) < 1e-9 * max(1.0, dref)
