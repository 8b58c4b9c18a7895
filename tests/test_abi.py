"""CPU tests: the C-ABI library loads, exports every symbol include/dsrg_b200.h declares, and
fails loudly (no CPU fallback) when no GPU is present.  No compute calls."""
import os
import re

import pytest

from conftest import ROOT
from dsrg_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "dsrg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsrg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_table_agree():
    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in header_functions():
        assert hasattr(L, name), name
    assert L.dsrg_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dsrg_b200 import api
    assert _lib.lib().dsrg_device_count() == 0
    with pytest.raises(_lib.DsrgError):
        api.Engine(1, 8, 8, 21)
    with pytest.raises(_lib.DsrgError):
        api.DenseCRF(8, 8, 21)


def test_product_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "dsrg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "liboracle" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_crf_params_default_follow_crf_py():
    from dsrg_b200 import api
    p = api.crf_params(scale_factor=12.0, color_factor=13, maxiter=10)
    assert (p.w1, p.w2, p.n_iters) == (10.0, 3.0, 10)
    import numpy as np
    assert p.theta_alpha_x == np.float32(80 / 12.0) and p.theta_gamma_y == np.float32(3 / 12.0)
    assert p.theta_beta_g == 13.0
