"""ctypes binding of libcpnest_b200.so (include/cpnest_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises ImportError with
the build instruction, and creating an engine without a CUDA device raises CpnError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CPNEST_B200_LIB: an alternative build of the same library (kernel experiments, scripts/build_variant.py)
LIB_PATH = os.environ.get("CPNEST_B200_LIB") or os.path.join(_HERE, "lib", "libcpnest_b200.so")


class CpnError(RuntimeError):
    pass


class cpn_config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("dim", C.c_int32), ("nlive", C.c_int32), ("maxmcmc", C.c_int32),
        ("nmcmc_init", C.c_int32), ("functor", C.c_int32), ("lanes", C.c_int32), ("ns_mode", C.c_int32),
        ("seed", C.c_uint64), ("nmcmc_safety", C.c_double), ("adapt_nmcmc", C.c_int32),
        ("world_size", C.c_int32), ("rank", C.c_int32), ("sampler", C.c_int32), ("rho_target", C.c_double),
        ("mix", C.c_int32 * 3), ("flags", C.c_int32),
    ]


class cpn_state(C.Structure):
    _fields_ = [
        ("logZ", C.c_double), ("info", C.c_double), ("logw", C.c_double), ("logLmin", C.c_double),
        ("logLmax", C.c_double), ("condition", C.c_double), ("nmcmc_exact", C.c_double), ("rho_max", C.c_double),
        ("iteration", C.c_int64), ("n_hist", C.c_int64), ("batches", C.c_int64),
        ("total_steps", C.c_int64), ("total_accepted", C.c_int64), ("total_evals", C.c_int64),
        ("failed_chains", C.c_int64), ("nmcmc", C.c_int32), ("done", C.c_int32),
        ("slice_mu", C.c_double), ("hmc_dt", C.c_double),
        ("rho_kind", C.c_double * 3), ("nmcmc_kind", C.c_int32 * 3), ("chains_kind", C.c_int32 * 3),
        ("two_lag_kind", C.c_int32 * 3), ("rho_decay_kind", C.c_double * 3),
    ]

    def as_dict(self):
        d = {}
        for n, _ in self._fields_:
            v = getattr(self, n)
            d[n] = list(v) if hasattr(v, "__len__") else v
        return d


FUNCTORS = dict(gaussian_iid=0, gaussian_corr=1, eggbox=2, rosenbrock=3, ackley=4, ring=5,
                gaussian_mean_sigma=6, gaussian_mixture=7, user=8, gaussian_mixture_nd=9)
WALK, STRETCH, DE, EIGEN = 0, 1, 2, 3
NS_SEQUENTIAL, NS_REFERENCE = 0, 1
SAMPLERS = dict(mh=0, slice=1, hmc=2)
SLICE_DIFF, SLICE_GAUSS, SLICE_CORR = 0, 1, 2

_P = C.POINTER
_dp = _P(C.c_double)
_ip = _P(C.c_int32)
_vp = C.c_void_p

# every symbol include/cpnest_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cpn_last_error": (C.c_char_p, []),
    "cpn_device_count": (C.c_int, []),
    "cpn_create": (C.c_int, [_P(cpn_config), _dp, _dp, _vp, C.c_size_t, _P(_vp)]),
    "cpn_destroy": (None, [_vp]),
    "cpn_set_cycle": (C.c_int, [_vp, _P(C.c_uint8), C.c_int32]),
    "cpn_set_slice_cycle": (C.c_int, [_vp, _P(C.c_uint8), C.c_int32]),
    "cpn_set_slice_mu": (C.c_int, [_vp, C.c_double]),
    "cpn_set_user_functor": (C.c_int, [_vp, C.c_char_p, _P(C.c_char_p), C.c_int32]),
    "cpn_set_seed": (C.c_int, [_vp, C.c_uint64]),
    "cpn_init_prior": (C.c_int, [_vp]),
    "cpn_set_live": (C.c_int, [_vp, _dp, C.c_int32]),
    "cpn_update_ensemble_stats": (C.c_int, [_vp]),
    "cpn_get_ensemble_stats": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "cpn_select_and_accumulate": (C.c_int, [_vp, C.c_int32]),
    "cpn_evolve_batch": (C.c_int, [_vp, C.c_int32]),
    "cpn_insert_batch": (C.c_int, [_vp, C.c_int32]),
    "cpn_staging_bytes": (C.c_int64, [_vp]),
    "cpn_set_staging": (C.c_int, [_vp, _P(_vp), C.c_int32]),
    "cpn_shard_record_doubles": (C.c_int64, [_vp]),
    "cpn_shard_range": (C.c_int, [_vp, C.c_int32, C.c_int32, _ip, _ip]),
    "cpn_kind_ranges": (C.c_int, [_ip, C.c_int32, _ip]),
    "cpn_pack_shard": (C.c_int, [_vp, C.c_int32, _vp]),
    "cpn_unpack_shards": (C.c_int, [_vp, C.c_int32, _vp]),
    "cpn_ns_step": (C.c_int, [_vp, C.c_int32]),
    "cpn_run": (C.c_int, [_vp, C.c_int32, C.c_int64, C.c_double, _P(C.c_int64)]),
    "cpn_finalise": (C.c_int, [_vp, _dp, _dp]),
    "cpn_checkpoint_bytes": (C.c_int64, [_vp]),
    "cpn_save_checkpoint": (C.c_int, [_vp, _vp, C.c_int64]),
    "cpn_load_checkpoint": (C.c_int, [_vp, _vp, C.c_int64]),
    "cpn_synchronize": (C.c_int, [_vp]),
    "cpn_get_state": (C.c_int, [_vp, _P(cpn_state)]),
    "cpn_set_nmcmc": (C.c_int, [_vp, C.c_int32]),
    "cpn_set_tolerance": (C.c_int, [_vp, C.c_double]),
    "cpn_get_live": (C.c_int, [_vp, _dp]),
    "cpn_num_dead": (C.c_int64, [_vp]),
    "cpn_get_dead": (C.c_int, [_vp, C.c_int64, C.c_int64, _dp, _dp]),
    "cpn_num_insertion_indices": (C.c_int64, [_vp]),
    "cpn_get_insertion_indices": (C.c_int, [_vp, C.c_int64, C.c_int64, _dp]),
    "cpn_get_last_batch": (C.c_int, [_vp, C.c_int32, _dp, _ip, _ip]),
    "cpn_evolve_host": (C.c_int, [_vp, C.c_int32, C.c_double, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "cpn_evolve_host_table": (C.c_int, [_vp, C.c_int32, C.c_double, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_int32]),
    "cpn_host_gather_rows": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp]),
    "cpn_host_scatter_rows": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp]),
    "cpn_set_stream": (C.c_int, [_vp, _vp]),
    "cpn_enable_chain_log": (C.c_int, [_vp, C.c_int32, C.c_int32]),
    "cpn_get_chain_log": (C.c_int, [_vp, C.c_int32, _dp, C.c_int32, _ip]),
    "cpn_launch_count": (C.c_int64, [_vp]),
    "cpn_pick_lanes": (C.c_int, [_vp, C.c_int32, _ip, _ip]),
    "cpn_eval_model": (C.c_int, [_vp, _dp, C.c_int32, _dp, _dp]),
    "cpn_eval_gradients": (C.c_int, [_vp, _dp, C.c_int32, _dp, _dp]),
    "cpn_replay_mh": (C.c_int, [_vp, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp, C.c_int32, _dp, C.c_int32,
                                C.c_int32, C.c_double, C.c_int32, C.c_int32, _dp, _ip, _ip, _dp]),
    "cpn_trace_mh": (C.c_int, [_vp, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                               C.c_int32, C.c_uint64, _dp, _ip, _ip, _ip, _dp]),
    "cpn_replay_slice": (C.c_int, [_vp, C.c_int32, _dp, C.c_int32, _dp, C.c_int32, _dp, _P(C.c_int64), C.c_double,
                                   C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, _dp, _ip, _ip, _ip, _dp]),
    "cpn_trace_slice": (C.c_int, [_vp, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp, _dp, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                  C.c_double, C.c_int32, C.c_uint64, _dp, _ip, _ip, _ip, _dp, _dp, C.c_int32]),
    "cpn_trace_hmc": (C.c_int, [_vp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_uint64,
                                _dp, _ip, _ip, _ip, _dp, C.c_int32]),
    "cpn_replay_hmc": (C.c_int, [_vp, C.c_int32, _dp, _dp, C.c_double, _dp, _P(C.c_int64), C.c_double, C.c_int32, C.c_int32,
                                 C.c_double, C.c_int32, _dp, _ip, _ip, _ip, _dp]),
    "cpn_slice_directions": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_double, _dp]),
    "cpn_stage_replacements": (C.c_int, [_vp, C.c_int32, _dp]),
    "cpn_ns_reset": (C.c_int, [_vp]),
    "cpn_ns_increment": (C.c_int, [_vp, _dp, C.c_int32, C.c_int32, C.c_int32]),
    "cpn_ns_trapezoid": (C.c_int, [_vp, _dp]),
    "cpn_philox": (C.c_int, [_vp, _P(C.c_uint32), _P(C.c_uint32), C.c_int32, _P(C.c_uint32)]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "cpnest_b200: CUDA library %s is missing. Build it with `python -m cpnest_b200.build` "
            "(needs nvcc; there is no CPU fallback)." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CpnError(load().cpn_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
