"""C-ABI surface: the shared library loads, exports every symbol include/potus_b200.h declares, the ctypes
mirrors match the header's struct layouts, and the product path fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "potus_b200.h")).read()
    return re.findall(r"POTUS_API\s+[\w\s\*]+?\b(potus_\w+)\s*\(", hdr)


def test_library_exports_every_declared_symbol(cuda_lib):
    from us_potus_model_b200 import cabi
    names = _declared_functions()
    assert len(names) == len(cabi.EXPORTS) >= 11 and set(names) == set(cabi.EXPORTS)
    for n in names:
        assert getattr(cuda_lib, n) is not None


def test_struct_sizes_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    import subprocess
    from us_potus_model_b200 import cabi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "potus_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(PotusData),sizeof(PotusConfig),sizeof(PotusStats),offsetof(PotusData,state_covariance_0),'
                   'offsetof(PotusConfig,seed),offsetof(PotusStats,n_params));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(cabi.PotusData), C.sizeof(cabi.PotusConfig), C.sizeof(cabi.PotusStats),
            cabi.PotusData.state_covariance_0.offset, cabi.PotusConfig.seed.offset, cabi.PotusStats.n_params.offset]
    assert got == want


def test_num_params_without_gpu(cuda_lib, datalists):
    from us_potus_model_b200 import cabi
    for y, D in ((2016, 15098), (2012, 14220), (2008, 14000)):
        pd, keep = cabi.marshal_data(datalists[y])
        assert cuda_lib.potus_num_params(C.byref(pd)) == D


def test_no_cpu_fallback(cuda_lib, pkg, datalists):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is only observable on a CPU-only box")
    from us_potus_model_b200 import cabi
    with pytest.raises(cabi.PotusError, match="no CUDA device|no CPU fallback|CUDA"):
        pkg.cmdstan_model().sample(data=datalists[2008], chains=2, iter_warmup=5, iter_sampling=5)
    with pytest.raises(cabi.PotusError):
        pkg.logp_grad(datalists[2008], np.zeros((1, 14000)))


def test_invalid_data_is_rejected_like_stan(cuda_lib, datalists):
    """Data-block constraints (poll_model_2020.stan:9-23,37) are re-checked before any device work."""
    from us_potus_model_b200 import cabi
    cfg = cabi.make_config(chains=1, iter_warmup=1, iter_sampling=1)

    def create(d):
        pd, keep = cabi.marshal_data(d)
        h = C.c_void_p()
        rc = cuda_lib.potus_create(C.byref(pd), C.byref(cfg), C.byref(h))
        msg = cuda_lib.potus_last_error().decode()
        if rc == 0:
            cuda_lib.potus_destroy(h)
        return rc, msg

    base = datalists[2016]
    d = dict(base); d["day_state"] = base["day_state"].copy(); d["day_state"][3] = 300
    rc, msg = create(d)
    assert rc == -1 and "day_state[4]" in msg and "[1, T]" in msg
    d = dict(base); d["poll_mode_national"] = base["poll_mode_national"].copy(); d["poll_mode_national"][0] = 9
    assert create(d)[0] == -1
    d = dict(base); cov = base["state_covariance_0"].copy(); cov[0, 1] += 1e-3; d["state_covariance_0"] = cov
    rc, msg = create(d)
    assert rc == -1 and "not symmetric" in msg
    d = dict(base); d["state_covariance_0"] = -base["state_covariance_0"]
    rc, msg = create(d)
    assert rc == -1
    d = dict(base); d["n_democrat_state"] = base["n_democrat_state"].copy(); d["n_democrat_state"][0] = 10 ** 6
    assert create(d)[0] == -1
    # sizes neither kernel family holds are refused with POTUS_ERR_UNSUPPORTED, not silently mangled
    # (S=60 is beyond the resident kernel and is routed to the streaming family: tests/test_gpu_stream.py)
    big = __import__("potus_pkg").load().synthetic_datalist(S=300, T=100, N_state=500, N_national=100, P=20)
    rc, msg = create(big)
    assert rc == -2 and "outside the streaming kernel" in msg
    big = __import__("potus_pkg").load().synthetic_datalist(S=60, T=600, N_state=500, N_national=100, P=20)
    assert create(big)[0] == -2


def test_hot_kernels_are_tcgen05_and_bulk_copy_code(cuda_lib):
    """The product kernels must be the sm_100a code the design describes, not a recompiled mma.sync / cp.async path: every
    sampler / evaluation kernel holds tcgen05 MMAs (SASS UTCHMMA), TMEM loads (LDTM), tcgen05.commit (UTCBAR) and 1-D bulk
    async copies (UBLKCP); nothing in the library uses the warp-level HMMA pipe.  (B200_PROFILING.md: PTX -> SASS names.)"""
    import shutil
    import subprocess
    from us_potus_model_b200 import build as b
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", b.LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : (\w+)\n", sass)
    fun = {parts[i]: parts[i + 1] for i in range(1, len(parts) - 1, 2)}
    for k in ("potus_nuts_kernel", "potus_eval_kernel", "potus_stream_kernel", "potus_stream_eval_kernel"):
        assert k in fun, sorted(fun)
        for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP"):
            assert re.search(rf"\b{mnemonic}\b", fun[k]), (k, mnemonic)
    assert "arch = sm_100a" in sass
    assert not re.search(r"\bHMMA\.", sass) and "LDGSTS" not in sass
