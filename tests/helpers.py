"""Shared helpers of the parity tests (inputs identical to tests/golden/make_golden.py)."""
import hashlib
import importlib.util
import os

import numpy as np

from dsrg_b200 import synth

_spec = importlib.util.spec_from_file_location(
    "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def renorm64(q_nchw):
    """pylayers.py:328-330 on a float32 (N,C,H,W) array of raw CRF marginals."""
    from oracle import crf_oracle
    return crf_oracle.renormalise(q_nchw)   # the reference's layout decides the float64 summation order


def srg_case_inputs(name):
    for n, H, W, cues, index, tweak in make_golden.SRG_CASES:
        if n == name:
            return make_golden.srg_inputs(H, W, cues, index, tweak)
    raise KeyError(name)


def crf_case_inputs(name):
    for i, (n, H, W, var, sf, kind) in enumerate(make_golden.CRF_CASES):
        if n == name:
            im, unary = make_golden.crf_inputs(H, W, var, kind, i)
            return im, unary, sf
    raise KeyError(name)
