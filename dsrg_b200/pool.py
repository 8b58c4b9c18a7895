"""One batch-1 engine per label count, shared by the per-image callers (krahenbuhl2013.CRF, the
inference post-processing): the evaluation tools feed images of many different sizes
(training/tools/test-ms.py:86-87), so the engine is sized for the largest image seen so far and
re-shaped per call (dsrg_engine_set_size) instead of being re-created."""
from . import api as _api

_ENGINES = {}      # (M, device) -> Engine
_ROUND = 64        # capacity granularity, so that slightly larger images do not force a new engine


def engine_for(H, W, M, device=None):
    """device None: the calling thread's current CUDA device (dsrg_current_device)."""
    H, W, M = int(H), int(W), int(M)
    if device is None:
        from . import _lib
        device = _lib.lib().dsrg_current_device()
    key = (M, int(device))
    eng = _ENGINES.get(key)
    if eng is not None:
        hc, wc = eng.capacity
        if H > hc or W > wc:
            eng.close()
            eng = None
            H0, W0 = max(H, hc), max(W, wc)
        else:
            H0, W0 = hc, wc
    else:
        H0, W0 = H, W
    if eng is None:
        up = lambda v: (v + _ROUND - 1) // _ROUND * _ROUND  # noqa: E731
        eng = _api.Engine(1, up(H0), up(W0), M, device)
        _ENGINES[key] = eng
    eng.set_size(H, W)
    return eng


def clear():
    for eng in _ENGINES.values():
        eng.close()
    _ENGINES.clear()
