import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CONFIGS = ["r0_00g0_99", "r0_01g0_99", "r0_05g0_95", "r0_00g0_9999"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


class Golden:
    """One decoded data/*.jld2 of the reference (tools/make_golden.py)."""

    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.walls = np.asfortranarray(d["walls"])
        self.H, self.W = self.walls.shape
        self.goal = tuple(int(v) - 1 for v in d["goal"])          # 0-based
        self.s0 = tuple(int(v) - 1 for v in d["s0"])
        self.s0_idx = self.s0[0] + self.H * self.s0[1]
        self.params = dict(randomness=float(d["randomness"]), step_cost=float(d["step_cost"]),
                           reward=float(d["reward"]), gamma=float(d["gamma"]))
        self.betas = d["betas"]
        self._states = d["states"] - 1
        self.nu = self.dense(d["nu"])
        self.traj = d["traj"]
        self.gap = d["gap"]

    def dense(self, v):
        a = np.full((self.H, self.W), np.nan, order="F")
        a[self._states[:, 0], self._states[:, 1]] = v
        return a

    def V(self, b, k):
        return self.dense(self.traj[b, k])

    def G(self, b, k):
        return self.dense(self.gap[b, k])

    def oracle(self):
        import oracle
        return oracle.Grid(self.walls, self.goal, **self.params)


@pytest.fixture(scope="session", params=CONFIGS)
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def golden01():
    return Golden("r0_01g0_99")


@pytest.fixture(scope="session")
def piql():
    import piql_jl_b200
    return piql_jl_b200


@pytest.fixture(scope="session")
def engine(piql):
    eng = piql.Engine(0)
    yield eng
    eng.close()


def random_gridworld(size, density, seed):
    """make_gridworld with a fixed numpy stream; returns (walls F-order uint8, goal)."""
    import piql_jl_b200 as piql
    gw = piql.make_gridworld(size, density=density, rng=np.random.default_rng(seed))
    return np.asfortranarray(gw.walls, dtype=np.uint8), tuple(gw.goal)
