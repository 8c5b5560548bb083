"""Edge cases of the hot path on the GPU against the oracle: the smallest legal grids, empty and ragged path batches,
start lists longer than the batch, path ids beyond 2^32 (both Philox counter words), a grid just above the single-CTA limit,
and the argument errors around them."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

P = dict(randomness=0.1, step_cost=0.02, reward=0.2, gamma=0.9)


def corridor(n):
    """3 x (n+2) wall array with one free row of n cells; goal at its far end."""
    w = np.ones((3, n + 2), dtype=np.uint8, order="F")
    w[1, 1:n + 1] = 0
    return w, (1, n)


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def test_single_free_cell_grid(engine):
    """3 x 3 array whose only free cell is the goal: every recursion is its terminal value, every path has length 0."""
    w = np.ones((3, 3), dtype=np.uint8, order="F")
    w[1, 1] = 0
    engine.load_grid(w, (1, 1), **P)
    nu, st = engine.solve_nu()
    assert nu[1, 1] == 0.0 and np.isnan(nu).sum() == 8 and st.converged
    V = np.zeros((3, 3), order="F")
    V[1, 1] = 0.3
    lz, _ = engine.solve_logz(V, beta=0.5)
    assert abs(lz[1, 1] + 0.3) < 1e-12                      # logZ(goal) = -V(goal)   value_iteration.jl:93-94
    gap, _ = engine.solve_gap(V, np.nan_to_num(nu))
    assert gap[1, 1] == 0.0
    sa, cnt, ls, mst = engine.mc_logz(V, beta=0.5, npaths=1000, seed=1)
    assert mst.transitions == 0 and int(cnt.sum()) == 1000
    assert abs(ls[1, 1] + 0.3) < 1e-6                       # fp32 record of -V(goal)
    logw, meta = engine.mc_records(1000)
    org = oracle.Grid(w, (1, 1), **P)
    lw_ref, meta_ref = org.rollout(org.policy_table(V, beta=0.5), 0.3, [4], 1000, seed=1)
    assert np.array_equal(meta, meta_ref) and np.array_equal(bits(logw), bits(lw_ref))


@pytest.mark.parametrize("n", [2, 5, 40])
def test_corridor_grids(engine, n):
    w, goal = corridor(n)
    engine.load_grid(w, goal, **P)
    org = oracle.Grid(w, goal, **P)
    nu, _ = engine.solve_nu()
    nu_ref, _ = org.solve("nu")
    assert np.array_equal(np.isnan(nu), np.isnan(nu_ref)) and np.nanmax(np.abs(nu - nu_ref)) < 1e-12
    V = 0.5 * np.nan_to_num(nu)
    v1, lz, _ = engine.update_v(V, beta=0.7, want_logz=True)
    assert np.nanmax(np.abs(v1 - org.logz_updateV(V, beta=0.7))) < 1e-11
    gap, _ = engine.solve_gap(V, np.nan_to_num(nu))
    gap_ref, _ = org.solve("gap", V=V, nu=np.nan_to_num(nu_ref))
    assert np.nanmax(np.abs(gap - gap_ref)) < 1e-11
    # rollouts from the far end, path ids beyond 2^32 so that the high Philox counter word is exercised too
    s0, path0, npaths = 1 + 3 * 1, (1 << 33) + 5, 4096 + 37
    table = engine.policy_table(V, beta=0.7)
    ref = org.policy_table(V, beta=0.7)      # device libm vs glibc: records agree to an ulp, wall masks exactly
    assert np.array_equal(bits(table[:, 3]), bits(ref[:, 3]))
    ulp = np.abs(table.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    ulp[:, 3] = 0                            # (differences of cancelling terms, ~1e-17, are compared absolutely)
    bad = ~((ulp <= 1) | (np.abs(table - ref) < 1e-12))
    assert not bad.any(), (np.argwhere(bad).tolist(), table[bad].tolist(), ref[bad].tolist())
    engine.mc_reset(starts=[s0])
    engine.mc_rollout(npaths, seed=3, starts=[s0], path0=path0)
    sa, cnt, ls, mst = engine.mc_finalize()
    logw, meta = engine.mc_records(npaths)
    lw_ref, meta_ref = org.rollout(table, float(V[goal]), [s0], npaths, seed=3, path0=path0)
    assert np.array_equal(meta, meta_ref) and np.array_equal(bits(logw), bits(lw_ref))
    assert mst.transitions == int((meta_ref >> 2).sum()) and int(cnt.sum()) == npaths
    mx, sm, c, tr = oracle.reduce_paths(lw_ref, meta_ref, 1, path0=path0)
    ref_sa, ref_s = oracle.logz_from_stats(mx, sm, c)
    ok = c[0] > 0
    assert tr == mst.transitions and np.array_equal(cnt[s0], c[0])
    assert np.allclose(sa[s0][ok], ref_sa[0][ok], rtol=1e-5, atol=1e-6) and np.isnan(sa[s0][~ok]).all()
    assert abs(ls.ravel(order="F")[s0] - ref_s[0]) < 1e-5


def test_empty_and_ragged_batches(engine, golden01):
    g = golden01
    engine.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    nfree = int((g.walls == 0).sum())
    # no paths at all: nothing sampled, every bucket empty
    sa, cnt, ls, mst = engine.mc_logz(V, beta=1.0, npaths=0, seed=0)
    assert mst.transitions == 0 and mst.paths == 0 and int(cnt.sum()) == 0
    assert np.isnan(sa).all() and np.isnan(ls).all()
    # fewer paths than start slots (paths cycle over the free cells: the first 7 get one path each)
    sa, cnt, ls, mst = engine.mc_logz(V, beta=1.0, npaths=7, seed=0)
    assert int(cnt.sum()) == 7 and int((cnt.sum(axis=1) > 0).sum()) == 7
    assert int(np.isfinite(ls).sum()) == 7
    free = np.flatnonzero(g.walls.ravel(order="F") == 0)
    assert np.array_equal(np.flatnonzero(cnt.sum(axis=1)), free[:7])
    # a batch that is not a multiple of the warp, CTA or start-list size; a shard that starts mid-cycle
    org = g.oracle()
    table = engine.policy_table(V, beta=1.0)
    starts = [int(free[3]), int(free[50]), int(free[120])]
    for npaths, path0 in ((1, 0), (33, 2), (257, nfree + 1), (1000, 5)):
        engine.mc_reset(starts=starts)
        engine.mc_rollout(npaths, seed=11, starts=starts, path0=path0)
        sa, cnt, ls, mst = engine.mc_finalize()
        logw, meta = engine.mc_records(npaths)
        lw_ref, meta_ref = org.rollout(table, float(V[g.goal]), starts, npaths, seed=11, path0=path0)
        assert np.array_equal(meta, meta_ref) and np.array_equal(bits(logw), bits(lw_ref)), (npaths, path0)
        mx, sm, c, tr = oracle.reduce_paths(lw_ref, meta_ref, 3, path0=path0)
        assert np.array_equal(cnt[starts], c) and tr == mst.transitions, (npaths, path0)
    # the same paths in two ragged shards accumulate to the same counts as one call
    engine.mc_reset(starts=starts)
    engine.mc_rollout(1000, seed=11, starts=starts, path0=5)
    one = engine.mc_finalize()
    engine.mc_reset(starts=starts)
    engine.mc_rollout(333, seed=11, starts=starts, path0=5)
    engine.mc_rollout(667, seed=11, starts=starts, path0=5 + 333)
    two = engine.mc_finalize()
    assert np.array_equal(one[1], two[1]) and one[3].transitions == two[3].transitions
    assert np.allclose(one[0][starts], two[0][starts], rtol=1e-5, atol=1e-6, equal_nan=True)


def test_horizon_one_and_truncation(engine, golden01):
    g = golden01
    engine.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    org = g.oracle()
    table = engine.policy_table(V, beta=1.0)
    for est in ("A", "B"):
        for hor in (1, 3):
            engine.mc_reset(starts=[g.s0_idx])
            engine.mc_rollout(512, seed=2, starts=[g.s0_idx], estimator=est, horizon=hor)
            engine.mc_finalize()
            logw, meta = engine.mc_records(512)
            lw_ref, meta_ref = org.rollout(table, float(V[g.goal]), [g.s0_idx], 512, seed=2, estimator=est, horizon=hor)
            assert np.array_equal(meta, meta_ref) and (meta >> 2).max() <= hor
            if est == "A":
                assert np.array_equal(bits(logw), bits(lw_ref))
            else:
                assert np.allclose(logw, lw_ref, rtol=1e-5, atol=1e-6)


def test_grid_just_above_the_single_cta_limit(engine):
    """8192 cells is the resident solver's limit: 91 x 91 = 8281 cells takes the tiled kernel with one tile row and a
    handful of tile columns, 90 x 91 = 8190 the single-CTA solver; both must match the oracle."""
    from conftest import random_gridworld
    for size in ((89, 89), (88, 89)):
        w, goal = random_gridworld(size, 0.2, 21)
        params = dict(randomness=0.05, step_cost=0.02, reward=0.2, gamma=0.95)
        engine.load_grid(w, goal, **params)
        org = oracle.Grid(w, goal, **params)
        nu, st = engine.solve_nu(tol=1e-11)
        assert st.kernel == (2 if w.size > 8192 else 1), (w.shape, st.kernel)
        nu_ref, _ = org.solve("nu", tol=1e-11)
        assert np.array_equal(np.isnan(nu), np.isnan(nu_ref))
        assert np.nanmax(np.abs(nu - nu_ref)) < 1e-9


def test_argument_errors_around_the_sampler(piql, golden01):
    K = piql._capi
    g = golden01
    eng = piql.Engine(0)
    eng.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    wall_cell = int(np.flatnonzero(g.walls.ravel(order="F") != 0)[25])
    for kw in (dict(starts=[wall_cell]), dict(starts=[-1]), dict(starts=[]), dict(horizon=0), dict(horizon=1 << 31),
               dict(npaths=-1)):
        args = dict(npaths=16)
        args.update(kw)
        with pytest.raises(piql.PiqlError) as e:
            eng.mc_logz(V, beta=1.0, **args)
        assert e.value.code == K.ERR_ARG, kw
    Vbad = V.copy()
    Vbad[g.goal] = np.nan
    with pytest.raises(piql.PiqlError) as e:
        eng.mc_logz(Vbad, beta=1.0, npaths=16)
    assert e.value.code == K.ERR_ARG
    with pytest.raises(piql.PiqlError) as e:
        eng.mc_records(1 << 30)
    assert e.value.code == K.ERR_ARG
    # the context survives all of that
    sa, cnt, ls, st = eng.mc_logz(V, beta=1.0, npaths=64, starts=[g.s0_idx])
    assert int(cnt.sum()) == 64
    eng.close()


def test_async_finalize_matches_the_blocking_call(engine, golden01):
    """piql_mc_finalize_async: the snapshot is finalised on the copy stream while the context's stream already runs the next
    iteration (whose reset clears the live accumulators); float64/uint64 and float32/uint32 outputs, and the pipelined
    ShardedEstimator, must reproduce piql_mc_finalize of the same rollouts."""
    from piql_jl_b200 import distributed as pd
    g = golden01
    engine.load_grid(g.walls, g.goal, **g.params)
    V = np.nan_to_num(g.V(7, 0))
    n = g.H * g.W
    ref = []
    for seed in (5, 6, 7):
        engine.policy_table(V, beta=1.0, download=False)
        engine.mc_reset()
        engine.mc_rollout(1 << 16, seed=seed)
        sa, cnt, ls, st = engine.mc_finalize()
        ref.append((sa.copy(), cnt.copy(), ls.copy(), int(st.transitions)))
    for fdt, cdt in ((np.float64, np.uint64), (np.float32, np.uint32)):
        est = pd.ShardedEstimator(engine, 0, 1, 1 << 16, out_dtype=fdt)
        got = []
        for seed in (5, 6, 7):
            prev = est.step(V, 1.0, seed=seed)
            if prev is not None:
                got.append((tuple(a.copy() for a in prev[0]), prev[1]))     # the two buffer sets alternate: copy before reuse
        got.append(est.finish())
        assert est.finish() is None
        for (sa, cnt, ls), st in got:
            assert sa.dtype == fdt and cnt.dtype == cdt and sa.shape == (n, 4)
        for ((sa, cnt, ls), st), (rsa, rcnt, rls, rtr) in zip(got, ref):
            assert int(st.transitions) == rtr and np.array_equal(cnt.astype(np.uint64), rcnt)
            ok = rcnt > 0
            tol = 0 if fdt == np.float64 else 1e-6
            assert np.allclose(sa[ok], rsa[ok], rtol=tol, atol=tol) and np.all(np.isnan(sa[~ok]))
            free = ~np.isnan(rls)
            assert np.allclose(ls[free], rls[free], rtol=tol, atol=tol)
