"""ctypes mirror of include/potus_b200.h and loader for the in-tree CUDA library.

The product path has NO CPU fallback: if lib/libpotus_b200.so is missing or fails to load this
module raises, and every compute entry point fails loudly when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpotus_b200.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class PotusData(C.Structure):
    _fields_ = [
        ("N_national_polls", C.c_int32), ("N_state_polls", C.c_int32), ("T", C.c_int32), ("S", C.c_int32),
        ("P", C.c_int32), ("M", C.c_int32), ("Pop", C.c_int32),
        ("state", _i32p), ("day_state", _i32p), ("day_national", _i32p), ("poll_state", _i32p), ("poll_national", _i32p),
        ("poll_mode_state", _i32p), ("poll_mode_national", _i32p), ("poll_pop_state", _i32p), ("poll_pop_national", _i32p),
        ("n_democrat_national", _i32p), ("n_two_share_national", _i32p), ("n_democrat_state", _i32p),
        ("n_two_share_state", _i32p),
        ("unadjusted_national", _f64p), ("unadjusted_state", _f64p), ("mu_b_prior", _f64p), ("state_weights", _f64p),
        ("sigma_c", C.c_double), ("sigma_m", C.c_double), ("sigma_pop", C.c_double),
        ("sigma_measure_noise_national", C.c_double), ("sigma_measure_noise_state", C.c_double),
        ("sigma_e_bias", C.c_double),
        ("state_covariance_0", _f64p),
        ("random_walk_scale", C.c_double), ("mu_b_T_scale", C.c_double), ("polling_bias_scale", C.c_double),
    ]


class PotusConfig(C.Structure):
    _fields_ = [
        ("chains", C.c_int32), ("chain_id_offset", C.c_int32), ("iter_warmup", C.c_int32), ("iter_sampling", C.c_int32),
        ("keep_per_chain", C.c_int32), ("max_treedepth", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
        ("seed", C.c_uint64), ("adapt_delta", C.c_double), ("init_radius", C.c_double),
        ("n_gpus", C.c_int32), ("reserved", C.c_int32),
    ]


class PotusStats(C.Structure):
    _fields_ = [
        ("n_leapfrog_total", C.c_int64), ("n_leapfrog_sampling", C.c_int64), ("n_divergent_sampling", C.c_int64),
        ("gpu_launches", C.c_int64),
        ("seconds_total", C.c_double), ("seconds_warmup", C.c_double), ("seconds_sampling", C.c_double),
        ("mean_stepsize", C.c_double), ("mean_accept_stat", C.c_double), ("mean_treedepth", C.c_double),
        ("n_params", C.c_int32), ("n_draws_kept", C.c_int32), ("seconds_gather", C.c_double),
    ]


_INT_VECS = ("state", "day_state", "day_national", "poll_state", "poll_national", "poll_mode_state",
             "poll_mode_national", "poll_pop_state", "poll_pop_national", "n_democrat_national",
             "n_two_share_national", "n_democrat_state", "n_two_share_state")
_DBL_VECS = ("unadjusted_national", "unadjusted_state", "mu_b_prior", "state_weights")
_SCALARS_I = ("N_national_polls", "N_state_polls", "T", "S", "P", "M", "Pop")
_SCALARS_D = ("sigma_c", "sigma_m", "sigma_pop", "sigma_measure_noise_national", "sigma_measure_noise_state",
              "sigma_e_bias", "random_walk_scale", "mu_b_T_scale", "polling_bias_scale")


def marshal_data(data: dict):
    """Named list (dict) -> (PotusData, keepalive).  Mirrors what the R .Call shim does with the
    SEXP list: look up by name, ignore unknown names, accept integral doubles for integer fields
    (R hands `state`, `poll_*`, `n_democrat_*` over as REALSXP; final_2016.R:436-460)."""
    keep = []
    pd = PotusData()
    for k in _SCALARS_I:
        v = data.get(k, 1 if k in ("M", "Pop") else None)
        if v is None:
            raise KeyError(f"data list is missing '{k}'")
        if float(v) != int(v):
            raise ValueError(f"{k} must be integral, got {v}")
        setattr(pd, k, int(v))
    for k in _SCALARS_D:
        default = {"sigma_m": 0.0, "sigma_pop": 0.0, "sigma_e_bias": 0.0}.get(k)
        v = data.get(k, default)
        if v is None:
            raise KeyError(f"data list is missing '{k}'")
        setattr(pd, k, float(v))
    full = "poll_mode_state" in data

    def _expect_len(name: str, arr) -> None:
        # Stan's own message shape for a short/long vector ("mismatch in dimension declared and found in context")
        want = pd.N_state_polls if name.endswith("_state") or name == "state" else (
            pd.N_national_polls if name.endswith("_national") else pd.S)
        if arr.ndim != 1 or arr.shape[0] != want:
            raise ValueError(f"Exception: mismatch in dimension declared and found in context; processing stage=data initialization; "
                             f"variable name={name}; position=0; dims declared=({want}); dims found=({','.join(map(str, arr.shape))})")

    for k in _INT_VECS:
        if k not in data:
            if k.startswith("poll_mode") or k.startswith("poll_pop"):
                setattr(pd, k, None)
                continue
            raise KeyError(f"data list is missing '{k}'")
        a = np.asarray(data[k])
        if a.dtype.kind == "f":
            if not np.all(a == np.rint(a)):
                raise ValueError(f"{k} must hold integers")
        arr = np.ascontiguousarray(a, dtype=np.int32)
        _expect_len(k, arr)
        keep.append(arr)
        setattr(pd, k, arr.ctypes.data_as(_i32p))
    for k in _DBL_VECS:
        if k not in data:
            if k.startswith("unadjusted") and not full:
                setattr(pd, k, None)
                continue
            raise KeyError(f"data list is missing '{k}'")
        arr = np.ascontiguousarray(np.asarray(data[k], dtype=np.float64))
        _expect_len(k, arr)
        keep.append(arr)
        setattr(pd, k, arr.ctypes.data_as(_f64p))
    cov = np.asarray(data["state_covariance_0"], dtype=np.float64)
    if cov.shape != (pd.S, pd.S):
        raise ValueError(f"Exception: mismatch in dimension declared and found in context; processing stage=data initialization; "
                         f"variable name=state_covariance_0; position=0; dims declared=({pd.S},{pd.S}); dims found=({','.join(map(str, cov.shape))})")
    cov = np.asfortranarray(cov)
    flat = np.ascontiguousarray(cov.reshape(-1, order="F"))
    keep.append(flat)
    pd.state_covariance_0 = flat.ctypes.data_as(_f64p)
    return pd, keep


def make_config(chains=4, iter_warmup=500, iter_sampling=500, seed=1843, keep_per_chain=0, max_treedepth=10,
                adapt_delta=0.8, init_radius=2.0, device=0, chain_id_offset=0, force_stream=False, n_gpus=1) -> PotusConfig:
    c = PotusConfig()
    c.chains, c.chain_id_offset, c.iter_warmup, c.iter_sampling = int(chains), int(chain_id_offset), int(iter_warmup), int(iter_sampling)
    c.keep_per_chain, c.max_treedepth, c.device, c.flags = int(keep_per_chain), int(max_treedepth), int(device), (1 if force_stream else 0)
    c.seed, c.adapt_delta, c.init_radius = int(seed), float(adapt_delta), float(init_radius)
    c.n_gpus, c.reserved = int(n_gpus), 0
    return c


_lib = None

EXPORTS = ("potus_create", "potus_run", "potus_run_begin", "potus_run_poll", "potus_run_end", "potus_set_state", "potus_draws_size", "potus_get_draws", "potus_get_stats",
           "potus_device_buffer", "potus_postprocess", "potus_destroy", "potus_last_error", "potus_logp_grad", "potus_logp_grad_ex", "potus_num_params")


def load_library(path: str | None = None):
    """dlopen the in-tree CUDA library and declare prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback for the product path.")
    lib = C.CDLL(p)
    lib.potus_create.argtypes = [C.POINTER(PotusData), C.POINTER(PotusConfig), C.POINTER(C.c_void_p)]
    lib.potus_create.restype = C.c_int
    lib.potus_run.argtypes = [C.c_void_p]
    lib.potus_run.restype = C.c_int
    lib.potus_run_begin.argtypes = [C.c_void_p]
    lib.potus_run_begin.restype = C.c_int
    lib.potus_run_poll.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.potus_run_poll.restype = C.c_int
    lib.potus_run_end.argtypes = [C.c_void_p]
    lib.potus_run_end.restype = C.c_int
    lib.potus_set_state.argtypes = [C.c_void_p, _f64p, _f64p, _f64p]
    lib.potus_set_state.restype = C.c_int
    lib.potus_draws_size.argtypes = [C.c_void_p, C.c_char_p]
    lib.potus_draws_size.restype = C.c_size_t
    lib.potus_get_draws.argtypes = [C.c_void_p, C.c_char_p, _f64p, C.c_size_t]
    lib.potus_get_draws.restype = C.c_int
    lib.potus_get_stats.argtypes = [C.c_void_p, C.POINTER(PotusStats)]
    lib.potus_get_stats.restype = C.c_int
    lib.potus_device_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.potus_device_buffer.restype = C.c_int
    lib.potus_postprocess.argtypes = [C.c_void_p, _f64p, C.c_double, _f64p, _f64p]
    lib.potus_postprocess.restype = C.c_int
    lib.potus_destroy.argtypes = [C.c_void_p]
    lib.potus_destroy.restype = None
    lib.potus_last_error.argtypes = []
    lib.potus_last_error.restype = C.c_char_p
    lib.potus_logp_grad.argtypes = [C.POINTER(PotusData), _f64p, C.c_int, _f64p, _f64p]
    lib.potus_logp_grad.restype = C.c_int
    lib.potus_logp_grad_ex.argtypes = [C.POINTER(PotusData), _f64p, C.c_int, _f64p, _f64p, C.c_int]
    lib.potus_logp_grad_ex.restype = C.c_int
    lib.potus_num_params.argtypes = [C.POINTER(PotusData)]
    lib.potus_num_params.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


class PotusError(RuntimeError):
    pass


def check(lib, rc: int):
    if rc != 0:
        msg = lib.potus_last_error()
        raise PotusError(f"potus_b200 error {rc}: {msg.decode() if msg else '?'}")
