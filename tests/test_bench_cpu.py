"""bench.py's host-side helpers (no GPU): the oracle-side slot statistics behind the bench line's `parity` object and the
bounded-sample start list of the CPU arm must agree with a plain reduction of all the step's paths."""
import os
import sys

import numpy as np

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _side():
    import bench
    walls, goal, params = bench.make_world("grid256")
    org = oracle.Grid(walls, goal, **params)
    V, _ = org.nu_jacobi_dense(50, reward_scale=bench.VLO_SCALE, nthreads=2)
    V = np.nan_to_num(V)
    table = org.policy_table(V, beta=bench.BETA, nthreads=2)
    return bench, bench.OracleSide(None, walls, goal, params, V, threads=2, table=table)


def test_slot_stats_equal_the_reduction_of_the_whole_step():
    bench, osd = _side()
    nfree = len(osd.free)
    id0, id1 = 5 * nfree + 17, 5 * nfree + 17 + 3 * nfree + 100      # a step that starts mid-cycle: 3-4 paths per slot
    lw, me = osd.records(id0, id1 - id0)
    mx, sm, cn, _tr = oracle.reduce_paths(lw, me, nfree, path0=id0)
    lsa, _ = oracle.logz_from_stats(mx, sm, cn)
    slots = np.array([0, 1, 16, 17, 18, nfree // 2, nfree - 1])
    c2, l2, npaths = osd.slot_stats(slots, id0, id1)
    assert np.array_equal(c2, cn[slots]) and npaths == int(cn[slots].sum())
    ok = c2 > 0
    assert np.allclose(l2[ok], lsa[slots][ok], rtol=0, atol=1e-12)


def test_sample_starts_reproduce_the_steps_paths():
    bench, osd = _side()
    nfree = len(osd.free)
    for path0, sample in ((0, 1000), (nfree - 300, 1000), (7 * nfree + 5, nfree + 10)):
        starts = bench.sample_starts(osd.free, path0, sample)
        a = osd.org.rollout(osd.table, osd.vgoal, starts, sample, path0=path0, seed=bench.MC_SEED)
        b = osd.org.rollout(osd.table, osd.vgoal, osd.free, sample, path0=path0, seed=bench.MC_SEED)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        assert len(starts) == min(sample, nfree)
