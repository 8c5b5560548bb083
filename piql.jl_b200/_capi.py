"""ctypes binding of include/piql_b200.h -- the same C ABI a Julia host reaches with `ccall`.

No compute happens in Python and there is no CPU fallback: if the shared library is missing
or no B200 is visible the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PIQL_B200_LIB names another build of the same library (the Julia shim honours the same variable)
LIB_PATH = os.environ.get("PIQL_B200_LIB") or os.path.join(_HERE, "lib", "libpiql_b200.so")

OK, ERR_ARG, ERR_STATE, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NOT_CONVERGED = range(7)
F64, F32, F32_FAST = 0, 1, 2
PRECISIONS = {"f64": F64, "f32": F32, "f32fast": F32_FAST, F64: F64, F32: F32, F32_FAST: F32_FAST}
EST_SAMPLED_STOP, EST_DISCOUNTED = 0, 1
START_ALL_FREE, START_LIST = 0, 1
KIND_NU, KIND_LOGZ, KIND_GAP, KIND_Z = 0, 1, 2, 3


class GridParams(C.Structure):
    _fields_ = [("randomness", C.c_double), ("step_cost", C.c_double), ("reward", C.c_double),
                ("gamma", C.c_double)]


class SolveOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_sweeps", C.c_int64), ("precision", C.c_int32),
                ("check_every", C.c_int32)]


class SolveStats(C.Structure):
    _fields_ = [("sweeps", C.c_int64), ("residual", C.c_double), ("device_seconds", C.c_double),
                ("backups", C.c_int64), ("converged", C.c_int32), ("kernel", C.c_int32)]


class McOpts(C.Structure):
    _fields_ = [("npaths", C.c_int64), ("path0", C.c_int64), ("seed", C.c_uint64), ("estimator", C.c_int32),
                ("start_mode", C.c_int32), ("horizon", C.c_int64), ("starts", C.POINTER(C.c_int32)),
                ("nstarts", C.c_int64), ("injected", C.POINTER(C.c_float)), ("injected_steps", C.c_int64)]


class McStats(C.Structure):
    _fields_ = [("transitions", C.c_uint64), ("paths", C.c_uint64), ("device_seconds", C.c_double),
                ("rollout_seconds", C.c_double)]


class TableProblem(C.Structure):
    _fields_ = [("nstates", C.c_int64), ("nactions", C.c_int32), ("row_ptr", C.POINTER(C.c_int64)),
                ("succ", C.POINTER(C.c_int32)), ("prob", C.POINTER(C.c_double)), ("reward", C.POINTER(C.c_double)),
                ("log_prior", C.POINTER(C.c_double)), ("terminal", C.POINTER(C.c_uint8)), ("gamma", C.c_double)]


class ViShardHandles(C.Structure):
    _fields_ = [("x0", C.c_uint8 * 64), ("x1", C.c_uint8 * 64), ("flags", C.c_uint8 * 64), ("np", C.c_int64),
                ("tj", C.c_int32), ("tiles_j", C.c_int32), ("col0", C.c_int64), ("col1", C.c_int64)]


class PeerHandles(C.Structure):
    _fields_ = [("lse_max", C.c_uint8 * 64), ("lse_sum", C.c_uint8 * 64), ("count", C.c_uint8 * 64),
                ("flags", C.c_uint8 * 64), ("n", C.c_int64)]


# every symbol include/piql_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
SYMBOLS = {
    "piql_version": (C.c_int32, []),
    "piql_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "piql_create": (C.c_int32, [C.POINTER(_P), C.c_int32]),
    "piql_destroy": (C.c_int32, [_P]),
    "piql_last_error": (C.c_char_p, [_P]),
    "piql_stream": (C.c_int32, [_P, C.POINTER(_P)]),
    "piql_grid_load": (C.c_int32, [_P, C.POINTER(C.c_uint8), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                   C.POINTER(GridParams)]),
    "piql_grid_free_count": (C.c_int32, [_P, C.POINTER(C.c_int64)]),
    "piql_solve_nu": (C.c_int32, [_P, C.c_double, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_solve_logz": (C.c_int32, [_P, _D, _D, C.c_double, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_update_v": (C.c_int32, [_P, _D, C.c_double, C.POINTER(SolveOpts), _D, _D, C.POINTER(SolveStats)]),
    "piql_solve_gap": (C.c_int32, [_P, _D, _D, _D, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_solve_z": (C.c_int32, [_P, _D, _D, C.c_double, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_z_update_v": (C.c_int32, [_P, _D, C.c_double, C.POINTER(SolveOpts), _D, _D, C.POINTER(SolveStats)]),
    "piql_sweep_once": (C.c_int32, [_P, C.c_int32, _D, _D, _D, C.c_double, C.c_double, C.c_int32, _D, _D, _D]),
    "piql_policy_table": (C.c_int32, [_P, _D, _D, C.c_double, C.POINTER(C.c_float)]),
    "piql_policy_table_set": (C.c_int32, [_P, C.POINTER(C.c_float), C.c_double]),
    "piql_mc_logz": (C.c_int32, [_P, _D, _D, C.c_double, C.POINTER(McOpts), _D, C.POINTER(C.c_uint64), _D,
                                 C.POINTER(McStats)]),
    "piql_trace_paths": (C.c_int32, [_P, C.POINTER(McOpts), C.c_int64, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "piql_path_estimates": (C.c_int32, [_P, _D, _D, C.POINTER(McOpts), C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "piql_mc_reset_dev": (C.c_int32, [_P, C.POINTER(McOpts)]),
    "piql_mc_rollout_dev": (C.c_int32, [_P, C.POINTER(McOpts)]),
    "piql_mc_stats_dev": (C.c_int32, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64)]),
    "piql_mc_rescale_dev": (C.c_int32, [_P, _P]),
    "piql_mc_finalize": (C.c_int32, [_P, _D, C.POINTER(C.c_uint64), _D, C.POINTER(McStats)]),
    "piql_mc_update_v": (C.c_int32, [_P, _D, C.c_double, C.POINTER(McOpts), C.c_int32, _D, C.POINTER(McStats)]),
    "piql_peer_export": (C.c_int32, [_P, C.POINTER(PeerHandles)]),
    "piql_peer_import": (C.c_int32, [_P, C.c_int32, C.c_int32, C.POINTER(PeerHandles)]),
    "piql_mc_allreduce_peer": (C.c_int32, [_P]),
    "piql_peer_close": (C.c_int32, [_P]),
    "piql_mc_timing": (C.c_int32, [_P, _D, _D, C.POINTER(C.c_int64)]),
    "piql_mc_transitions_total": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "piql_policy_table16": (C.c_int32, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "piql_policy_table_async": (C.c_int32, [_P, _D, _D, C.c_double]),
    "piql_mc_finalize_async": (C.c_int32, [_P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "piql_mc_finalize_wait": (C.c_int32, [_P, C.POINTER(McStats)]),
    "piql_mc_records": (C.c_int32, [_P, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int64]),
    "piql_vi_prepare_dev": (C.c_int32, [_P, C.c_int32, _D, _D, _D, C.c_double, C.c_double, C.c_int32]),
    "piql_learning_traj": (C.c_int32, [_P, _D, _D, C.c_int32, C.c_int32, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_solve_gap_batch": (C.c_int32, [_P, _D, C.c_int32, _D, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_table_load": (C.c_int32, [_P, C.POINTER(TableProblem)]),
    "piql_table_set_q": (C.c_int32, [_P, _D]),
    "piql_table_solve": (C.c_int32, [_P, C.c_int32, _D, _D, C.c_double, C.c_double, C.POINTER(SolveOpts), _D, C.POINTER(SolveStats)]),
    "piql_table_update_v": (C.c_int32, [_P, _D, C.c_double, C.POINTER(SolveOpts), _D, _D, C.POINTER(SolveStats)]),
    "piql_table_mc_logz": (C.c_int32, [_P, _D, C.c_double, C.POINTER(McOpts), _D, C.POINTER(C.c_uint64), _D, C.POINTER(McStats),
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "piql_vi_shard_begin": (C.c_int32, [_P, C.c_int32, C.c_int32, C.POINTER(ViShardHandles)]),
    "piql_vi_shard_connect": (C.c_int32, [_P, C.POINTER(ViShardHandles)]),
    "piql_vi_shard_end": (C.c_int32, [_P]),
    "piql_vi_shard_wait_stats": (C.c_int32, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "piql_vi_set_iterate": (C.c_int32, [_P, C.POINTER(C.c_double)]),
    "piql_vi_tile_plan": (C.c_int32, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "piql_vi_sweeps_dev": (C.c_int32, [_P, C.c_int64]),
    "piql_vi_fetch": (C.c_int32, [_P, _D, _D]),
    "piql_launch_count": (C.c_int32, [_P, C.POINTER(C.c_int64)]),
}

_lib = None


class PiqlError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("piql_b200 error %d: %s" % (code, message))
        self.code = code


def load():
    """dlopen the in-tree shared library and type every entry point.  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
