"""ctypes access to the C oracle (oracle/potus_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SO = os.path.join(_HERE, "_build", "libpotus_oracle.so")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import potus_pkg  # noqa: E402

cabi = potus_pkg.load().cabi if hasattr(potus_pkg.load(), "cabi") else None
if cabi is None:
    import importlib
    cabi = importlib.import_module("us_potus_model_b200.cabi")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "potus_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-fvisibility=hidden", "-o", SO, src,
                               "-lm", "-lpthread"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        f64p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orc_model_create.argtypes = [C.POINTER(cabi.PotusData)]
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_num_params.argtypes = [C.c_void_p]
        L.orc_num_params.restype = C.c_int
        L.orc_logp_grad.argtypes = [C.c_void_p, f64p, f64p, C.c_int]
        L.orc_logp_grad.restype = C.c_double
        L.orc_constrain.argtypes = [C.c_void_p] + [f64p] * 7
        L.orc_sample.argtypes = [C.c_void_p, C.POINTER(cabi.PotusConfig), C.c_int, C.c_int, C.c_int, C.c_int, f64p, f64p, f64p,
                                 f64p, i64p, f64p]
        L.orc_sample.restype = C.c_double
        L.orc_transitions.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double, f64p, f64p, C.c_uint32,
                                      C.c_int, f64p, f64p, C.c_int]
        L.orc_rng_words.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleModel:
    def __init__(self, data: dict):
        self.data = data
        self._pd, self._keep = cabi.marshal_data(data)
        self.h = lib().orc_model_create(C.byref(self._pd))
        if not self.h:
            raise ValueError("state_covariance_0 is not positive definite")
        self.D = lib().orc_num_params(self.h)
        self.S, self.T = int(data["S"]), int(data["T"])

    def __del__(self):
        try:
            if self.h:
                lib().orc_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def logp_grad(self, theta, literal=False):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        g = np.empty(self.D)
        lp = lib().orc_logp_grad(self.h, _p(th), _p(g), int(literal))
        return lp, g

    def constrain(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        d = self.data
        out = dict(mu_b=np.empty(self.S * self.T), mu_c=np.empty(int(d["P"])), mu_m=np.zeros(int(d.get("M", 1))),
                   mu_pop=np.zeros(int(d.get("Pop", 1))), e_bias=np.zeros(self.T), polling_bias=np.empty(self.S))
        lib().orc_constrain(self.h, _p(th), _p(out["mu_b"]), _p(out["mu_c"]), _p(out["mu_m"]), _p(out["mu_pop"]),
                            _p(out["e_bias"]), _p(out["polling_bias"]))
        out["mu_b"] = out["mu_b"].reshape(self.T, self.S).T.copy()  # [S,T]
        return out

    def sample(self, chains=4, iter_warmup=500, iter_sampling=500, seed=1843, threads=None, literal=False, tree_mode=0,
               max_iters=0, save_theta=False, chain_id_offset=0, max_treedepth=10, adapt_delta=0.8, save_inv_metric=False):
        cfg = cabi.make_config(chains=chains, iter_warmup=iter_warmup, iter_sampling=iter_sampling, seed=seed,
                               chain_id_offset=chain_id_offset, max_treedepth=max_treedepth, adapt_delta=adapt_delta)
        threads = threads or min(chains, os.cpu_count() or 1)
        theta = np.zeros((chains, iter_sampling, self.D)) if save_theta else None
        mon = np.zeros((chains, iter_sampling, self.S + 1))
        stats = np.full((chains, iter_warmup + iter_sampling, 7), np.nan)
        eps = np.zeros(chains)
        nlf = np.zeros(chains, dtype=np.int64)
        im = np.zeros((chains, self.D)) if save_inv_metric else None
        secs = lib().orc_sample(self.h, C.byref(cfg), int(literal), int(tree_mode), int(threads), int(max_iters),
                                _p(theta) if save_theta else None, _p(mon), _p(stats), _p(eps),
                                nlf.ctypes.data_as(C.POINTER(C.c_int64)), _p(im) if save_inv_metric else None)
        return dict(inv_metric=im, theta=theta, monitor=mon, stats=stats, stepsize=eps, n_leapfrog=nlf, seconds=secs, threads=threads)

    def transitions(self, q0, eps, inv_metric, n_iter=1, seed=1843, chain=0, tree_mode=1, max_depth=10, iter0=0, literal=False):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        im = np.ascontiguousarray(inv_metric, dtype=np.float64)
        q_out = np.empty((n_iter, self.D))
        stats = np.empty((n_iter, 7))
        lib().orc_transitions(self.h, int(seed), int(chain), int(tree_mode), int(max_depth), float(eps), _p(im), _p(q0),
                              int(iter0), int(n_iter), _p(q_out), _p(stats), int(literal))
        return q_out, stats


def rng_words(seed, chain, idx, it, stream, sub=0):
    out = (C.c_uint32 * 4)()
    lib().orc_rng_words(int(seed), int(chain), int(idx), int(it), int(stream), int(sub), out)
    return list(out)
