"""Host-side logic and the C-ABI surface, without a GPU: the reference-interface mirror
(ControlProblem / gridworld_ctrl / modify / Growable / tabulate), the shared library's exported
symbols versus include/piql_b200.h, and that nothing falls back to the CPU when no GPU exists."""
import ctypes
import math
import os
import re
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT, random_gridworld


def test_library_exports_every_declared_symbol(piql):
    hdr = open(os.path.join(ROOT, "include", "piql_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(piql_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(piql._capi.SYMBOLS), declared ^ set(piql._capi.SYMBOLS)
    lib = ctypes.CDLL(piql._capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert piql._capi.load().piql_version() == 100


def test_ctypes_structs_match_the_header_layout(piql):
    K = piql._capi
    assert ctypes.sizeof(K.GridParams) == 32
    assert ctypes.sizeof(K.SolveOpts) == 24 and K.SolveOpts.precision.offset == 16
    assert ctypes.sizeof(K.SolveStats) == 40
    assert ctypes.sizeof(K.McOpts) == 72 and K.McOpts.starts.offset == 40
    assert ctypes.sizeof(K.McStats) == 32


def test_no_cpu_fallback_without_a_gpu(piql):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(piql.PiqlError) as e:
        piql.Engine(0)
    assert e.value.code == piql._capi.ERR_NO_DEVICE
    gw = piql.make_gridworld((6, 6), density=0.1, rng=np.random.default_rng(0))
    with pytest.raises(piql.PiqlError):
        piql.generate_nu(piql.gridworld_ctrl(gw))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "piql.jl_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("like the oracle", ""), (f, "mentions oracle")


def test_make_gridworld_semantics(piql):
    gw = piql.make_gridworld((18, 18), density=0.2, rng=np.random.default_rng(1))
    w = gw.walls
    assert w.shape == (20, 20) and w[0].all() and w[-1].all() and w[:, 0].all() and w[:, -1].all()
    assert not w[gw.goal]                                          # test/runtests.jl:28
    walled, nreach = oracle.wall_unreachable(w, gw.goal[0] + 20 * gw.goal[1])
    assert np.array_equal(walled != 0, w) and nreach == (~w).sum()  # every free cell reaches the goal
    ctrl = piql.gridworld_ctrl(gw, randomness=0.1, gamma=0.99, rng=np.random.default_rng(2))
    for _ in range(20):
        assert not w[ctrl.initial_state()]                          # test/runtests.jl:29
    assert [piql.take_step((5, 5), a, np.zeros((9, 9), bool)) for a in range(4)] == [(6, 5), (4, 5), (5, 6), (5, 4)]
    assert len(piql.state_iterator(gw)) == (~w).sum() - 1


def test_closures_agree_with_the_oracle_algebra(piql):
    walls, goal = random_gridworld((8, 8), 0.15, seed=4)
    gw = piql.Gridworld(walls != 0, goal)
    ctrl = piql.gridworld_ctrl(gw, randomness=0.1, gamma=0.97)
    org = oracle.Grid(walls, goal, randomness=0.1, gamma=0.97)
    V, _ = org.solve("nu", reward_scale=0.8)
    Vf = piql.Growable.from_dense(V, walls)
    Q = lambda s, a: piql.controlV(ctrl, Vf, s, a)
    H = walls.shape[0]
    tab = org.policy_table(np.nan_to_num(V), beta=0.7)
    for s in piql.state_iterator(gw)[::5]:
        idx = s[0] + H * s[1]
        p = [piql.action_probability(ctrl, Q, s, a) for a in range(4)]
        assert abs(sum(p) - 1) < 1e-12
        assert np.allclose(np.cumsum(p)[:3], tab[idx, :3], atol=2e-7)
        assert np.allclose([0.7 * piql.fitness(ctrl, Vf, Q, s, a) for a in range(4)], tab[idx, 4:], atol=1e-7)
        assert abs(piql.kl_cost(ctrl, s, Q) - org.kl_cost(np.nan_to_num(V), idx)) < 1e-12
        assert piql.kl_cost(ctrl, s, Q) > 0                         # test/runtests.jl:39
    # the goal cell's own record (only path_transitions, which walks on through the goal, reads it): a blocked move out of
    # the goal stays in the goal and earns the reward -- reward_function looks at s' alone   GridWorld.jl:116
    w2 = np.ones((3, 4), dtype=np.uint8, order="F")
    w2[1, 1:3] = 0
    gw2 = piql.Gridworld(w2 != 0, (1, 2))
    ctrl2 = piql.gridworld_ctrl(gw2, randomness=0.1, gamma=0.9)
    org2 = oracle.Grid(w2, (1, 2), randomness=0.1, gamma=0.9)
    V2 = 0.5 * np.nan_to_num(org2.solve("nu")[0])
    Vf2 = piql.Growable.from_dense(np.where(w2 != 0, np.nan, V2), w2)
    Q2 = lambda s, a: piql.controlV(ctrl2, Vf2, s, a)
    tab2 = org2.policy_table(V2, beta=0.7)
    for s in ((1, 1), (1, 2)):
        idx = s[0] + 3 * s[1]
        p = [piql.action_probability(ctrl2, Q2, s, a) for a in range(4)]
        assert np.allclose(np.cumsum(p)[:3], tab2[idx, :3], atol=2e-7)
        assert np.allclose([0.7 * piql.fitness(ctrl2, Vf2, Q2, s, a) for a in range(4)], tab2[idx, 4:], atol=1e-7)
    # logpropagatorexp == log(propagator_average(exp f))
    f = lambda s: 0.3 * s[0] - 0.1 * s[1]
    s = piql.state_iterator(gw)[3]
    assert abs(ctrl.logpropagatorexp(s, 2, f) - math.log(ctrl.propagator_average(s, 2, lambda x: math.exp(f(x))))) < 1e-12


def test_tabulate_reads_modified_rewards(piql):
    gw = piql.make_gridworld((7, 7), density=0.1, rng=np.random.default_rng(3))
    ctrl = piql.gridworld_ctrl(gw, randomness=0.05, gamma=0.9)
    t = piql.tabulate(ctrl)
    assert (t.randomness, t.gamma) == (0.05, 0.9) and t.reward == 0.2 and t.step_cost == 0.1 * 0.2  # 0.020000000000000004, as stored in data/*.jld2
    weaker = piql.modify(ctrl, reward_function=lambda s, a, ss: 0.8 * ctrl.reward_function(s, a, ss))
    t2 = piql.tabulate(weaker)
    assert t2.reward == 0.8 * 0.2 and t2.step_cost == 0.8 * t.step_cost
    assert piql.modify(ctrl, γ=0.5).gamma == 0.5
    with pytest.raises(NotImplementedError):
        piql.tabulate(piql.modify(ctrl, reward_function=lambda s, a, ss: float(s[0])))
    with pytest.raises(NotImplementedError):
        piql.tabulate(piql.modify(ctrl, propagator_average=lambda s, a, f: 0.0))


def test_growable_grows_like_the_reference(piql):
    g = piql.Growable((2, 3), default=lambda s: 1.5)
    assert g.dict == {(2, 3): 0.0}
    assert g((4, 4)) == 1.5 and (4, 4) in g.dict and len(g.dict) == 2
    walls = np.ones((4, 4), bool); walls[1:3, 1:3] = False
    arr = np.arange(16.0).reshape(4, 4)
    d = piql.Growable.from_dense(arr, walls)
    assert d((1, 2)) == 6.0 and set(d.dict) == {(1, 1), (1, 2), (2, 1), (2, 2)}
    assert np.isnan(piql.Growable((1, 1)).dense(walls)).sum() == 15


def test_host_path_sampler_mirrors_reference_draw_order(piql):
    # choose_action consumes one uniform; the propagator one uniform and one integer draw, always
    class Counting:
        def __init__(self): self.r = np.random.default_rng(0); self.n_random = 0; self.n_int = 0
        def random(self): self.n_random += 1; return self.r.random()
        def integers(self, n): self.n_int += 1; return self.r.integers(n)
    gw = piql.make_gridworld((6, 6), density=0.0, rng=np.random.default_rng(0))
    rng = Counting()
    ctrl = piql.gridworld_ctrl(gw, randomness=0.3, rng=rng)
    V = piql.Growable.from_dense(np.zeros((8, 8)), gw.walls)
    Q = lambda s, a: piql.controlV(ctrl, V, s, a)
    steps = list(piql.path_transitions(ctrl, (3, 3), Q, rng, nsteps=20))
    assert len(steps) == 20 and rng.n_random == 40 and rng.n_int == 20
    assert len(piql.fitness_reward_estimate(ctrl, (3, 3), V, Q, rng, nsteps=20)) == 20


def test_reference_arm_of_the_bench_runs_on_cpu_and_keeps_the_contract():
    """bench.py --impl reference needs no GPU: one JSON line with the contract's keys from rank 0 (all host threads even
    when a launcher exported OMP_NUM_THREADS=1, as torchrun does), nothing from the other ranks."""
    import json
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "grid256", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--ref-paths", "8192", "--setup-sweeps", "20"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["n_gpus"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    env["RANK"] = "1"
    out1 = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_library_path_override_fails_loudly_when_missing(tmp_path):
    """PIQL_B200_LIB selects another build of the library (the Julia shim reads the same variable); a path that does not
    exist is an import error naming it, never a silent fallback to anything else."""
    import subprocess
    missing = str(tmp_path / "libpiql_nowhere.so")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import piql_jl_b200 as piql\n"
            "try:\n    piql._capi.load()\nexcept ImportError as e:\n    print('IMPORTERROR', e)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PIQL_B200_LIB=missing), capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "IMPORTERROR" in r.stdout and missing in r.stdout, (r.stdout, r.stderr[-300:])


def test_grid_key_is_content_based(piql):
    """solvers.engine_for reloads the device grid iff the tabulated problem changed: the key hashes the wall bytes (tabulate()
    copies the wall array, so an address would never match; an in-place edit must not leave a stale grid on the device)."""
    from piql_jl_b200 import gridworld as G
    gw = piql.make_gridworld((12, 9), density=0.2, rng=np.random.default_rng(4))
    ctrl = piql.gridworld_ctrl(gw, randomness=0.05, gamma=0.97)
    k1, k2 = G.tabulate(ctrl).key(), G.tabulate(ctrl).key()
    assert k1 == k2 and hash(k1) == hash(k2)
    assert G.tabulate(piql.gridworld_ctrl(gw, randomness=0.06, gamma=0.97)).key() != k1
    i, j = np.argwhere(~gw.walls)[3]
    if (i, j) == tuple(gw.goal):
        i, j = np.argwhere(~gw.walls)[4]
    gw.walls[i, j] = True                      # in place: same object, same address, other problem
    assert G.tabulate(ctrl).key() != k1
