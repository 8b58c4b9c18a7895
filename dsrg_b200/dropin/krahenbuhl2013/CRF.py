"""Drop-in for CRF/krahenbuhl2013/CRF.py: same function name, arguments, defaults and return value,
computed by the batched B200 engine (a batch of one) instead of a per-call CPU DenseCRFWrapper."""
import numpy as np

from dsrg_b200 import api as _api
from dsrg_b200.pool import engine_for as _engine
from krahenbuhl2013.wrapper import DenseCRF  # noqa: F401  (re-exported like the reference module)

__all__ = ["CRF", "DenseCRF"]


def CRF(image, unary, maxiter=10, scale_factor=1.0, color_factor=13):
    """Mean-field inference in a fully connected CRF with Gaussian edge potentials.

    image : (H, W, 3) array with values in [0, 256); cast to ubyte exactly like CRF.py:32
    unary : (H, W, M) array; the unary ENERGY is ``-unary`` (CRF.py:28), i.e. pass log-probabilities
            (test-ms.py:106) or probabilities (pylayers.py:82) as the reference's callers do
    Returns the (H, W, M) float32 marginals after ``maxiter`` iterations with the reference's pairwise
    terms: bilateral w=10, sigma 80/scale_factor and color_factor; spatial w=3, sigma 3/scale_factor
    (CRF.py:31-32).
    """
    assert(image.shape[:2] == unary.shape[:2])
    H, W, M = unary.shape
    params = _api.crf_params(scale_factor, color_factor, maxiter)
    u = np.ascontiguousarray(unary, dtype=np.float32).reshape(1, H, W, M)
    im = np.ascontiguousarray(np.asarray(image).astype('ubyte')).reshape(1, H, W, 3)
    return _engine(H, W, M).crf_host(u, im, params)[0]
