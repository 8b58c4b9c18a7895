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


def test_wire_format_host_helpers_roundtrip():
    """The 1-bit-per-value PCIe wire format of the *_host entry points (csrc/wire.cu): pack -> unpack is
    the identity on 0/1 planes at every alignment, non-0/1 input is rejected, the clamp mask sets 1e-4."""
    import ctypes as C
    import numpy as np
    L = _lib.lib()
    rng = np.random.RandomState(0)
    for n in (1, 5, 31, 32, 33, 100, 21 * 41 * 41, 21 * 321 * 321):
        for off in range(4):
            buf = np.zeros(n + 8, np.float32)
            buf[off:off + n] = (rng.rand(n) < 0.2).astype(np.float32)
            bits = np.zeros((n + 31) // 32 + 1, np.uint32)
            ok = L.dsrg_wire_pack_mask(C.c_void_p(buf[off:].ctypes.data), C.c_void_p(bits.ctypes.data), n)
            assert ok == 1
            want_bits = np.packbits(buf[off:off + n].astype(bool), bitorder="little")
            assert np.array_equal(bits.view(np.uint8)[: want_bits.size], want_bits)
            out = np.full(n + 8, -7.0, np.float32)
            L.dsrg_wire_unpack_mask(C.c_void_p(bits.ctypes.data), C.c_void_p(out[off:].ctypes.data), n)
            assert np.array_equal(out[off:off + n], buf[off:off + n])
            assert out[off + n] == -7.0 and (off == 0 or out[off - 1] == -7.0)
    bad = np.zeros(64, np.float32)
    bad[10] = 0.5
    assert L.dsrg_wire_pack_mask(C.c_void_p(bad.ctypes.data), C.c_void_p(np.zeros(3, np.uint32).ctypes.data), 64) == 0
    p = np.full(70, 0.5, np.float32)
    m = np.zeros(3, np.uint32)
    m[0], m[2] = 1 << 3, 1 << 5
    L.dsrg_wire_apply_clamp_mask(C.c_void_p(m.ctypes.data), C.c_void_p(p.ctypes.data), 70)
    assert p[3] == np.float32(1e-4) and p[69] == np.float32(1e-4) and p[4] == 0.5
