"""ctypes binding of libb200dde.so (include/b200dde.h).

This is the same set of calls the Julia shim (delaydiffeq.jl_b200/julia/B200DDE.jl) makes with
`ccall`; Julia is not installed in this image, so this module is the executable host.  There is no
fallback: if the shared library is missing or no sm_100 GPU is present, calls raise.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("B200DDE_LIB") or os.path.join(PKG_DIR, "lib", "libb200dde.so")  # env override: kernel experiments

ALG_TSIT5, ALG_ROSENBROCK23, ALG_BS3 = 0, 1, 2
MODELS = {
    "bc_model": 0, "mackey_glass": 1, "state_dependent": 2, "stiff_kinetics": 3, "delayed_logistic": 4,
    "constant_1delay": 5, "constant_2delays": 6, "constant_1delay_long": 7, "constant_1delay_dependent": 8,
    "backwards": 9, "sir": 10, "sir_ddae": 11,
}
RETCODES = {0: "Default", 1: "Success", 2: "MaxIters", 3: "DtLessThanMin", 4: "Unstable", 5: "DtNaN",
            6: "HistoryOverflow", 7: "QueueOverflow", 8: "Terminated", 9: "InexactDivision"}
STAT_NAMES = ("naccept", "nreject", "nf", "nfpiter", "nfpconvfail", "ntracked", "maxnodes", "nsolve")
N_STATS = 8


class B200DDEError(RuntimeError):
    pass


class Config(C.Structure):
    """b2d_config, field for field."""
    _fields_ = [
        ("struct_size", C.c_int32), ("alg", C.c_int32), ("constrained", C.c_int32), ("fp_max_iter", C.c_int32),
        ("fp_kappa", C.c_double), ("fp_div_rule", C.c_int32), ("controller_pow", C.c_int32),
        ("model_id", C.c_int32), ("order_discontinuity_t0", C.c_int32), ("neutral", C.c_int32),
        ("n_state", C.c_int32), ("n_param", C.c_int32), ("n_const_lags", C.c_int32), ("n_dep_lags", C.c_int32),
        ("abstol", C.c_double), ("reltol", C.c_double), ("dt0", C.c_double), ("dtmin", C.c_double),
        ("dtmax", C.c_double), ("maxiters", C.c_int64),
        ("gamma", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double), ("qsteady_min", C.c_double),
        ("qsteady_max", C.c_double), ("qoldinit", C.c_double), ("failfactor", C.c_double),
        ("disc_interp_points", C.c_int32), ("track_theta_mode", C.c_int32),
        ("disc_abstol", C.c_double), ("disc_reltol", C.c_double),
        ("rb23_dT_mode", C.c_int32), ("ring_capacity", C.c_int32), ("device", C.c_int32), ("chunk_traj", C.c_int32),
        ("fp_alg", C.c_int32), ("fp_max_history", C.c_int32), ("fp_aa_start", C.c_int32), ("fp_anderson_reset", C.c_int32),
        ("callback", C.c_int32), ("cb_affect", C.c_int32), ("cb_idx", C.c_int32), ("cb_save_positions", C.c_int32),
        ("cb_par0", C.c_double), ("cb_par1", C.c_double), ("cb_par2", C.c_double), ("cb_par3", C.c_double),
        ("cb_cond_idx", C.c_int32), ("cb_direction", C.c_int32), ("cb_interp_points", C.c_int32), ("cb_reserved", C.c_int32),
    ]


_lib = None

_PROTOTYPES = {
    # name: (restype, argtypes)
    "b2d_config_default": (None, [C.POINTER(Config), C.c_int32]),
    "b2d_model_info": (C.c_int, [C.c_int32] + [C.POINTER(C.c_int32)] * 5),
    "b2d_model_mass_matrix": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32]),
    "b2d_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config)]),
    "b2d_destroy": (C.c_int, [C.c_void_p]),
    "b2d_set_tolerances": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "b2d_set_save_idxs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "b2d_set_tstops": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "b2d_set_trajectory_lags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "b2d_set_d_discontinuities": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "b2d_solve": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p]),
    "b2d_solve_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "b2d_solve_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "b2d_solve_everystep": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2d_solve_dense": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2d_create_user_model": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_char_p, C.c_char_p]),
    "b2d_compile_user_model": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]),
    "b2d_release_cached": (C.c_int, []),
    "b2d_host_alloc": (C.c_void_p, [C.c_uint64]),
    "b2d_host_free": (None, [C.c_void_p]),
    "b2d_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b2d_measure_fp64_tflops": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "b2d_comm_unique_id": (C.c_int, [C.c_void_p]),
    "b2d_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "b2d_gather_saveat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "b2d_eval_model": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 5),
    "b2d_check_division": (C.c_int, [C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b2d_last_error": (C.c_char_p, []),
    "b2d_version": (C.c_int, []),
}


def exported_symbols():
    return sorted(_PROTOTYPES)


def lib():
    """Load libb200dde.so; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200DDEError(
                f"{LIB_PATH} not found: build it with `make -C {PKG_DIR}` (or __graft_entry__.build()). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().b2d_last_error().decode()


def check(rc):
    if rc != 0:
        raise B200DDEError(last_error())
