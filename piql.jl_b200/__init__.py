"""piql.jl_b200 -- B200 (sm_100a) accelerator for PIQL.jl's Monte-Carlo path sampling and soft
value-iteration hot path, behind the reference's own problem/solver interface.

The directory name carries a dot, so import it through the `piql_jl_b200` shim at the repo root:

    import piql_jl_b200 as piql
    gw = piql.make_gridworld((18, 18), density=0.2)
    ctrl = piql.gridworld_ctrl(gw, randomness=0.01, gamma=0.99)
    nu = piql.generate_nu(ctrl)

Layout: csrc/ (CUDA kernels + the C ABI of include/piql_b200.h), lib/ (the built shared library),
_capi.py (ctypes binding), engine.py (one context per GPU), gridworld.py and solvers.py (the
host-side mirror of the reference interface), jld2.py (writer of the reference's result files), build.py.
"""
from . import _capi
from . import jld2
from ._capi import PiqlError
from .engine import Engine
from .gridworld import (ControlProblem, Gridworld, gridworld_ctrl, make_gridworld, make_walls, modify,
                        state_iterator, tabulate, take_step)
from .tables import (BanditState, ProblemTable, bandit_state_iterator, binary_bandit_problem, reachable, tabulate_problem,
                     update)
from .solvers import (Growable, McResult, Qnormalization, action_probability, average_reward, choose_action,
                      controlQ, controlV, direct_reward_estimate, engine_for, fitness, fitness_reward_estimate, generate_logz, generate_nu,
                      generate_optimality_gap, generate_optimality_gap_batch, generate_z, generate_ν, kl_cost, learning_traj, learning_traj_batch, logz_updateV, mc_logz, mc_logz_updateV,
                      path_estimates, path_transitions, run_experiment, run_experiment_distributed, z_updateV)
