"""oracle/loss_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of BalancedSeedLossLayer (pylayers/pylayers/pylayers.py:120-152).  The
reference builds the forward/backward with Theano 0.8.2 (absent here, un-vendored): forward is
restated expression by expression (:129-139), backward is the analytic gradient of that
expression (what ``T.grad`` returns).  Pinned against the reference's OWN layer classes: their class
bodies are executed in place by oracle/ref_layers.py (Theano itself is absent and replaced by
oracle/theano_shim.py, torch autograd underneath); tests/golden/layers_ref.npz freezes those outputs
and tests/test_oracle_golden.py holds these restatements to them at 1e-12 (float64 evaluation) -- also
for SoftmaxLayer and ConstrainLossLayer below.  What stays unpinned is Theano's own float32 reduction
order (unspecified), hence the relative tolerance (1e-5) of the GPU tests.
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


# ---- SoftmaxLayer (pylayers.py:23-51) and ConstrainLossLayer (pylayers.py:154-180) ----
def _softmax(preds):
    e = np.exp(preds - np.max(preds, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def softmax_layer_forward(preds, dtype=np.float64):
    """probs = softmax + min_prob, renormalised (pylayers.py:33-36)."""
    s = _softmax(np.asarray(preds, dtype))
    probs = s + MIN_PROB
    return probs / np.sum(probs, axis=1, keepdims=True)


def softmax_layer_backward(preds, top_diff, dtype=np.float64):
    """T.grad(sum(probs * top_diff), preds) (pylayers.py:38-41), analytically."""
    s = _softmax(np.asarray(preds, dtype))
    g = np.asarray(top_diff, dtype)
    z = np.sum(s + MIN_PROB, axis=1, keepdims=True)
    gs = g / z - np.sum(g * (s + MIN_PROB), axis=1, keepdims=True) / (z * z)
    return s * (gs - np.sum(gs * s, axis=1, keepdims=True))


def constrain_loss(probs, log_smooth, dtype=np.float64):
    """mean_{n,h,w} sum_c ps * log(clip(ps / probs, 0.05, 20)), ps = exp(log_smooth) (pylayers.py:163-165)."""
    p = np.asarray(probs, dtype)
    ps = np.exp(np.asarray(log_smooth, dtype))
    return float(np.mean(np.sum(ps * np.log(np.clip(ps / p, 0.05, 20)), axis=1)))


def constrain_loss_grad(probs, log_smooth, dtype=np.float64):
    """(d loss / d probs, d loss / d log_smooth) (pylayers.py:168, :176-180)."""
    p = np.asarray(probs, dtype)
    ps = np.exp(np.asarray(log_smooth, dtype))
    ratio = ps / p
    inside = ((ratio >= 0.05) & (ratio <= 20)).astype(dtype)
    cnt = float(p.shape[0] * p.shape[2] * p.shape[3])
    return -(ps / p) * inside / cnt, ps * (np.log(np.clip(ratio, 0.05, 20)) + inside) / cnt
