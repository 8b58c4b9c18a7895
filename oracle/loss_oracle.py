"""oracle/loss_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of BalancedSeedLossLayer (pylayers/pylayers/pylayers.py:120-152).  The
reference builds the forward/backward with Theano 0.8.2 (absent here, un-vendored): forward is
restated expression by expression (:129-139), backward is the analytic gradient of that
expression (what ``T.grad`` returns).  Parity unpinned (no Theano to run, no reference test).
float32 like ``T.ftensor4``; the reduction order inside Theano is unspecified, so tests use a
relative tolerance (1e-5) against a float64 evaluation of the same formula.
"""
import numpy as np

MIN_PROB = 0.0001  # pylayers.py:20


def balanced_seed_loss(probs, labels, dtype=np.float64):
    """probs, labels: (N,C,H,W).  Returns the scalar loss (pylayers.py:129-139)."""
    p = np.asarray(probs, dtype)
    l = np.asarray(labels, dtype)
    probs_bg, labels_bg = p[:, 0], l[:, 0]
    probs_fg, labels_fg = p[:, 1:], l[:, 1:]
    count_bg = labels_bg.sum(axis=(1, 2), keepdims=True)
    count_fg = labels_fg.sum(axis=(1, 2, 3), keepdims=True)
    loss_1 = -np.mean((labels_bg * np.log(probs_bg)).sum(axis=(1, 2), keepdims=True) / np.maximum(count_bg, MIN_PROB))
    loss_2 = -np.mean((labels_fg * np.log(probs_fg)).sum(axis=(1, 2, 3), keepdims=True) / np.maximum(count_fg, MIN_PROB))
    return loss_1 + loss_2


def balanced_seed_loss_grad(probs, labels, dtype=np.float64):
    """d loss / d probs (what pylayers.py:142,151 obtains from T.grad): -lab / (p * cnt * N)."""
    p = np.asarray(probs, dtype)
    l = np.asarray(labels, dtype)
    N = p.shape[0]
    g = np.zeros_like(p)
    count_bg = np.maximum(l[:, 0].sum(axis=(1, 2)), MIN_PROB)
    count_fg = np.maximum(l[:, 1:].sum(axis=(1, 2, 3)), MIN_PROB)
    g[:, 0] = -l[:, 0] / (p[:, 0] * count_bg[:, None, None] * N)
    g[:, 1:] = -l[:, 1:] / (p[:, 1:] * count_fg[:, None, None, None] * N)
    return g
