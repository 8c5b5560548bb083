"""Host-side mirror of PIQL's problem plugin API for the gridworld (the drop-in boundary).

Mirrors, with the same names and argument meaning:
    ControlProblem, modify            src/PIQL.jl:12-49
    Gridworld, make_gridworld         src/Problems/GridWorld.jl:35-38, 78-88
    gridworld_ctrl                    src/Problems/GridWorld.jl:90-121
    state_iterator                    src/Problems/GridWorld.jl:123

In the reference the solvers see a problem only through eight closures.  Closures cannot cross a
C ABI, so `tabulate` turns a ControlProblem back into the flat description piql_grid_load takes
(wall array, goal, randomness, step cost, reward, gamma) -- the same step the Julia shim performs
(INTEGRATION.md).  States are (i, j) tuples, 0-based indices into the wall array (border included).
"""
from dataclasses import dataclass, replace
from typing import Any, Callable

import numpy as np

ACTIONS = range(4)            # step_choices(dim=2) = 0:3                     GridWorld.jl:20-22
_STEP = ((1, 0), (-1, 0), (0, 1), (0, -1))  # step(a)                          GridWorld.jl:24-26


def step(a):
    return _STEP[a]


def take_step(x0, a, walls):
    """ifelse(walls[x1], x0, x1)   GridWorld.jl:29-32"""
    x1 = (x0[0] + _STEP[a][0], x0[1] + _STEP[a][1])
    return x0 if walls[x1] else x1


@dataclass(frozen=True)
class Gridworld:
    walls: np.ndarray  # bool (H, W), border included
    goal: tuple


@dataclass(frozen=True)
class ControlProblem:
    """The functions that completely define a KL-control problem (src/PIQL.jl:12-23)."""
    action_space: Any
    action_prior: Callable          # (s, a) -> probability
    propagator: Callable            # (s, a) -> random next state
    reward_function: Callable       # (s0, a, s1) -> reward, in entropic units
    propagator_average: Callable    # (s, a, f) -> E_{s'} f(s')
    logpropagatorexp: Callable      # (s, a, f) -> log E_{s'} exp f(s')
    terminal_condition: Callable    # s -> bool
    initial_state: Callable         # () -> s
    gamma: float
    # tabulation hints recorded by gridworld_ctrl (not part of the reference struct)
    grid: Any = None
    randomness: float = None


def modify(ctrl, **fields):
    """modify(ctrl; field = value, ...)   src/PIQL.jl:28-49 (γ is spelled gamma)."""
    if "γ" in fields:
        fields["gamma"] = fields.pop("γ")
    if "propagator_average" in fields or "logpropagatorexp" in fields or "propagator" in fields:
        fields.setdefault("grid", None)  # dynamics replaced: no longer the implicit gridworld
    return replace(ctrl, **fields)


def make_walls(size, density=0.1, rng=None):
    """Border walls plus i.i.d. barriers, one draw per cell in column-major order   GridWorld.jl:40-52"""
    rng = np.random.default_rng() if rng is None else rng
    H, W = size[0] + 2, size[1] + 2
    walls = np.zeros((H, W), dtype=bool, order="F")
    walls[0, :] = walls[-1, :] = True
    walls[:, 0] = walls[:, -1] = True
    draws = rng.random(H * W).reshape((H, W), order="F")
    walls |= draws < density
    return walls


def draw_not_wall(walls, rng):
    """Rejection-sample a free cell (the goal may be returned too)   GridWorld.jl:54-61"""
    H, W = walls.shape
    while True:
        ii = int(rng.integers(H * W))
        s = (ii % H, ii // H)
        if not walls[s]:
            return s


def get_reachable(walls, goal):
    """Mask of the cells connected to the goal through free cells   GridWorld.jl:63-76"""
    from scipy import ndimage
    lab, _ = ndimage.label(~walls, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    return lab == lab[goal]


def make_gridworld(size, density=0.1, rng=None):
    """make_gridworld(size; density)   GridWorld.jl:78-88: unreachable cells become walls."""
    rng = np.random.default_rng() if rng is None else rng
    walls = make_walls(size, density=density, rng=rng)
    goal = draw_not_wall(walls, rng)
    walls = np.asfortranarray(walls | ~get_reachable(walls, goal))
    return Gridworld(walls, goal)


def gridworld_ctrl(gw, randomness=0.1, reward_scale=0.2, step_cost=None, reward=None, gamma=0.99, γ=None, rng=None):
    """gridworld_ctrl(gw; randomness, reward_scale, step_cost, reward, γ)   GridWorld.jl:90-121"""
    if γ is not None:
        gamma = γ
    step_cost = 0.1 * reward_scale if step_cost is None else step_cost
    reward = 1 * reward_scale if reward is None else reward
    walls, goal = gw.walls, tuple(gw.goal)
    n = len(ACTIONS)
    rng = np.random.default_rng() if rng is None else rng

    def slip(a, aa):
        return (a == aa) * (1 - randomness) + randomness / n

    def pa(s, a, f):
        out = 0.0
        for aa in ACTIONS:
            out += slip(a, aa) * f(take_step(s, aa, walls))
        return out

    def log_pa(s, a, f):
        with np.errstate(divide="ignore"):
            x = np.array([np.log(slip(a, aa)) + f(take_step(s, aa, walls)) for aa in ACTIONS])
        m = x.max()
        return m if np.isinf(m) else m + np.log(np.exp(x - m).sum())

    def propagator(s, a):
        # rand() < randomness and sample(0:3) are both always drawn (eager ifelse)   GridWorld.jl:113-115
        u, alt = rng.random(), int(rng.integers(n))
        return take_step(s, alt if u < randomness else a, walls)

    return ControlProblem(
        action_space=ACTIONS,
        action_prior=lambda s, a: 1.0 / n,
        propagator=propagator,
        reward_function=lambda s0, a, s1: reward if s1 == goal else -step_cost,
        propagator_average=pa,
        logpropagatorexp=log_pa,
        terminal_condition=lambda s: s == goal,
        initial_state=lambda: draw_not_wall(walls, rng),
        gamma=float(gamma),
        grid=gw,
        randomness=float(randomness),
    )


def state_iterator(gw):
    """Free, non-goal cells in column-major order   GridWorld.jl:123"""
    H, W = gw.walls.shape
    return [(i, j) for j in range(W) for i in range(H) if not gw.walls[i, j] and (i, j) != tuple(gw.goal)]


@dataclass(frozen=True)
class GridTable:
    walls: np.ndarray   # uint8 (H, W) Fortran order
    goal: tuple
    randomness: float
    step_cost: float
    reward: float
    gamma: float

    def key(self):
        """Identity of the tabulated problem: a content hash of the wall bytes (not their address -- tabulate() copies the
        wall array on every call, and a caller may edit walls in place) plus goal and parameters.  An unchanged problem is
        loaded on the device once, a changed one always reloaded (solvers.engine_for)."""
        import hashlib
        digest = hashlib.blake2b(np.ascontiguousarray(self.walls.T).data, digest_size=16).hexdigest()
        return (digest, self.walls.shape, self.goal, self.randomness, self.step_cost, self.reward, self.gamma)


def tabulate(ctrl, probes=64):
    """ControlProblem -> the flat arguments of piql_grid_load.

    Works for problems whose dynamics are gridworld_ctrl's; the reward function may have been
    replaced through modify() (as test/runtests.jl:7-8 and scripts/experiments.jl:26 do) provided
    it still depends only on whether the next state is the goal, which is checked by probing it.
    """
    gw = ctrl.grid
    if gw is None:
        raise NotImplementedError("only gridworld_ctrl problems can be tabulated for the implicit-grid ABI "
                                  "(a sparse-table ABI for general ControlProblems is future work)")
    goal = tuple(gw.goal)
    walls = np.asfortranarray(gw.walls, dtype=np.uint8)
    free = np.argwhere(gw.walls == 0)
    other = [tuple(int(v) for v in c) for c in free[:: max(1, len(free) // probes)] if tuple(c) != goal]
    reward = float(ctrl.reward_function(goal, 0, goal))
    if not other:
        step_cost = 0.0
    else:
        vals = {float(ctrl.reward_function(c, a, c)) for c in other for a in (0, 3)}
        if len(vals) != 1:
            raise NotImplementedError("reward_function must depend only on (next state == goal)")
        step_cost = -vals.pop()
    for c in other[:4]:  # the goal reward must not depend on where the step came from
        if float(ctrl.reward_function(c, 1, goal)) != reward:
            raise NotImplementedError("reward_function must depend only on (next state == goal)")
    return GridTable(walls, goal, float(ctrl.randomness), step_cost, reward, float(ctrl.gamma))
