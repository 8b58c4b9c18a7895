"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the inference post-processing of the reference's
evaluation tools.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this module; the product (dsrg_b200/) never does.

Restates, statement by statement, what predict_mask() does after net.forward():
  * training/tools/test-ms.py:84-111         -> predict_mask_ms
  * training/tools/generate_train_gt.py:76-104 -> predict_mask_gt
with the same third-party calls the reference makes (scipy.ndimage.zoom order=1, numpy float32 exp /
log / argmax) and the CRF restatement of oracle/crf_oracle.py for krahenbuhl2013.CRF.

Parity status: UNPINNED by the reference (it ships no test or golden vector for these tools, SURVEY.md
8c); scipy is 1.18 here against the reference's pinned 0.18 (python-dependencies.txt:15) -- same
corner-aligned order-1 mapping (grid_mode=False).
"""
import numpy as np
from scipy import ndimage as nd

from . import crf_oracle

EPS = 0.00001


def _crf(im, unary):
    return crf_oracle.CRF(im, unary, scale_factor=1.0)


def predict_mask_ms(im, blobs, smooth=True):
    """test-ms.py:86-111.  `blobs`: net.blobs['fc8-SEC'].data[0] of each scale, (M,h,w) float32.
    Returns (result (H,W) int64, probs (H,W,M) float32 = CRF marginals or the clamped softmax)."""
    d1, d2 = float(im.shape[0]), float(im.shape[1])                                  # :87
    scores_all = 0                                                                   # :89
    for blob in blobs:                                                               # :90
        scores = np.transpose(blob, [1, 2, 0])                                       # :95
        scores = nd.zoom(scores, (d1 / scores.shape[0], d2 / scores.shape[1], 1.0), order=1)   # :96
        scores_all += scores                                                         # :97
    scores_exp = np.exp(scores_all - np.max(scores_all, axis=2, keepdims=True))      # :99
    probs = scores_exp / np.sum(scores_exp, axis=2, keepdims=True)                   # :100
    probs[probs < EPS] = EPS                                                         # :102-103
    if smooth:
        probs = _crf(im, np.log(probs))                                              # :106
    return np.argmax(probs, axis=2), probs                                           # :106 / :109


def predict_mask_gt(im, blob, labels, smooth=True):
    """generate_train_gt.py:78-102.  `labels`: the image tags (1-based class ids, without background)."""
    scores = np.transpose(blob, [1, 2, 0])                                           # :85
    d1, d2 = float(im.shape[0]), float(im.shape[1])                                  # :86
    scores_exp = np.exp(scores - np.max(scores, axis=2, keepdims=True))              # :88
    probs = scores_exp / np.sum(scores_exp, axis=2, keepdims=True)                   # :89
    probs = nd.zoom(probs, (d1 / probs.shape[0], d2 / probs.shape[1], 1.0), order=1)  # :90
    probs[probs < EPS] = EPS                                                         # :92-93
    if smooth:
        probs = _crf(im, np.log(probs))                                              # :96
    labels = list(np.asarray(labels).tolist())                                       # :98
    labels.insert(0, 0)                                                              # :99
    probs_selected = probs[:, :, labels]                                             # :100
    probs_c = np.argmax(probs_selected, axis=2)                                      # :101
    result = np.asarray(labels)[probs_c]                                             # :102 (np.vectorize of a lookup)
    return result, probs
