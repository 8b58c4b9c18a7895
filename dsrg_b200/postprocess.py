"""Inference post-processing of the reference's evaluation tools on the GPU: everything predict_mask()
does after net.forward() (training/tools/test-ms.py:84-111, training/tools/generate_train_gt.py:76-104).
The network forward pass stays with the caller; pass the fc8 score blobs as they come out of
``net.blobs['fc8-SEC'].data[0]`` ((M, h, w) float32).  No CPU fallback: needs the CUDA library."""
import numpy as np

from . import api as _api
from .pool import engine_for as _engine_for

EPS = 0.00001  # test-ms.py:103, generate_train_gt.py:90


def _blob(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 3:
        raise ValueError("score blob must be (M, h, w)")
    return a


def _image(im, smooth):
    im = np.asarray(im)
    if im.ndim != 3 or im.shape[2] != 3:
        raise ValueError("image must be (H, W, 3)")
    return np.ascontiguousarray(im.astype('ubyte')) if smooth else None  # CRF.py:32


def predict_mask_ms(im, scores_per_scale, smooth=True, return_probs=False):
    """test-ms.py:84-111 after the forward passes: sum of zoomed score maps -> softmax -> clamp ->
    CRF(im, log(probs), scale_factor=1.0) -> argmax.  `scores_per_scale`: the (M,h,w) blobs of the
    241/321/401 passes (any number).  Returns the (H,W) label map (int64 like np.argmax)."""
    blobs = [_blob(s) for s in scores_per_scale]
    H, W = np.asarray(im).shape[:2]
    eng = _engine_for(H, W, blobs[0].shape[0])
    out = eng.predict_mask_host(blobs, _image(im, smooth), _api.crf_params(1.0), _api.POST_SUM_SCORES, EPS, smooth,
                                None, return_probs)
    if return_probs:
        return out[0].astype(np.int64), out[1]
    return out.astype(np.int64)


def predict_mask_gt(im, scores, labels, smooth=True, return_probs=False):
    """generate_train_gt.py:76-104 after net.forward(): softmax at network resolution -> zoom -> clamp ->
    CRF -> argmax restricted to [0] + labels, mapped back to label ids."""
    blob = _blob(scores)
    H, W = np.asarray(im).shape[:2]
    sel = [0] + [int(v) for v in np.asarray(labels).tolist()]   # :96-97
    eng = _engine_for(H, W, blob.shape[0])
    out = eng.predict_mask_host([blob], _image(im, smooth), _api.crf_params(1.0), _api.POST_ZOOM_PROBS, EPS, smooth,
                                sel, return_probs)
    if return_probs:
        return out[0].astype(np.int64), out[1]
    return out.astype(np.int64)
