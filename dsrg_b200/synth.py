"""Synthetic inputs for the DSRG pixel-labelling pass (fixed generator).

Host-side numpy only; used by tests/, bench.py and tests/golden/make_golden.py.  Shapes follow
the reference's blobs (pylayers/pylayers/pylayers.py:297-299): labels (C,), probs (C,H,W) f32,
cues (C,H,W) f32 0/1, image (H,W,3) u8.
"""
import numpy as np

C_DEFAULT = 21


def _gauss_blur_1d(a, sigma, axis):
    r = max(1, int(3 * sigma + 0.5))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    a = np.moveaxis(a, axis, -1)
    pad = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(r, r)], mode="reflect")
    out = np.zeros_like(a, dtype=np.float64)
    for i, w in enumerate(k):
        out += w * pad[..., i:i + a.shape[-1]]
    return np.moveaxis(out, -1, axis)


def make_labels(rng, C=C_DEFAULT):
    labels = np.zeros(C, np.float32)
    labels[0] = 1
    k = rng.choice([1, 2, 3])
    labels[rng.choice(np.arange(1, C), size=k, replace=False)] = 1
    return labels


def make_probs(rng, labels, H, W, C=C_DEFAULT):
    logits = rng.randn(C, H, W)
    ys, xs = np.mgrid[0:H, 0:W]
    for c in np.where(labels == 1)[0]:
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        s = max(H, W) / 4.0
        logits[c] += 8.0 * np.exp(-0.5 * (((ys - cy) / s) ** 2 + ((xs - cx) / s) ** 2))
    logits -= logits.max(0, keepdims=True)
    e = np.exp(logits)
    return (e / e.sum(0, keepdims=True)).astype(np.float32)


def make_cues(rng, labels, probs, variant="cam"):
    C, H, W = probs.shape
    cues = np.zeros((C, H, W), np.float32)
    if variant == "cam":
        am = probs.argmax(0)
        for c in np.where(labels == 1)[0]:
            cues[c] = ((rng.rand(H, W) < 0.02) & (am == c)).astype(np.float32)
    elif variant == "random":      # multi-class seeds + seeds of absent classes
        cues[...] = (rng.rand(C, H, W) < 0.01).astype(np.float32)
    else:
        raise ValueError(variant)
    return cues


def make_image(rng, H, W, variant="smooth"):
    if variant == "smooth":
        im = rng.rand(H, W, 3)
        s = max(H, W) / 20.0
        im = _gauss_blur_1d(_gauss_blur_1d(im, s, 0), s, 1)
        lo, hi = im.min((0, 1), keepdims=True), im.max((0, 1), keepdims=True)
        return np.round((im - lo) / (hi - lo) * 255.0).astype(np.uint8)
    if variant == "noise":
        return rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    if variant == "photo":
        # natural-image statistics: amplitude spectrum ~ 1/f (power 1/f^2), a shared luminance field plus weaker
        # independent chroma fields, stretched to the full 8-bit range with a few percent clipped like a photo
        fy, fx = np.fft.fftfreq(H)[:, None], np.fft.fftfreq(W)[None, :]
        f = np.sqrt(fy * fy + fx * fx)
        f[0, 0] = 1.0
        amp = 1.0 / f
        amp[0, 0] = 0.0

        def field():
            ph = np.exp(2j * np.pi * rng.rand(H, W))
            x = np.real(np.fft.ifft2(amp * ph))
            return x / x.std()
        lum = field()
        im = np.stack([lum + 0.35 * field() for _ in range(3)], axis=-1)
        lo, hi = np.percentile(im, 1.0), np.percentile(im, 99.0)
        return np.round(np.clip((im - lo) / (hi - lo), 0.0, 1.0) * 255.0).astype(np.uint8)
    raise ValueError(variant)


def make_problem(index, H, W, C=C_DEFAULT, cues="cam", image="smooth", seed=1234):
    """One image's worth of inputs; ``rng = RandomState(seed + index)``."""
    rng = np.random.RandomState(seed + index)
    labels = make_labels(rng, C)
    probs = make_probs(rng, labels, H, W, C)
    cu = make_cues(rng, labels, probs, cues)
    im = make_image(rng, H, W, image)
    return dict(labels=labels, probs=probs, cues=cu, image=im)


def make_batch(B, H, W, C=C_DEFAULT, cues="cam", image="smooth", seed=1234, start=0, unique=None):
    """Batch of ``B`` problems.  ``unique`` (<= B) bounds how many distinct images are generated
    (the rest repeat cyclically) so that big bench batches stay cheap to synthesise."""
    u = B if unique is None else min(B, unique)
    ps = [make_problem(start + i, H, W, C, cues, image, seed) for i in range(u)]
    pick = [ps[i % u] for i in range(B)]
    return dict(labels=np.stack([p["labels"] for p in pick]),
                probs=np.stack([p["probs"] for p in pick]),
                cues=np.stack([p["cues"] for p in pick]),
                image=np.stack([p["image"] for p in pick]))


def make_score_blobs(index, H, W, sizes=(31, 41, 51), C=C_DEFAULT, image="smooth", seed=4321):
    """Inputs of the inference post-processing (training/tools/test-ms.py:84-111): one (C,h,w) float32
    score blob per network scale -- the same blobby scene seen at every scale plus independent noise --
    the (H,W,3) uint8 image and the image tags (class ids without background)."""
    rng = np.random.RandomState(seed + index)
    labels = make_labels(rng, C)
    centres = [(c, rng.uniform(0, 1), rng.uniform(0, 1)) for c in np.where(labels == 1)[0]]
    blobs = []
    for s in sizes:
        ys, xs = np.mgrid[0:s, 0:s] / float(max(s - 1, 1))
        sc = rng.randn(C, s, s)
        for c, cy, cx in centres:
            sc[c] += 6.0 * np.exp(-0.5 * (((ys - cy) / 0.25) ** 2 + ((xs - cx) / 0.25) ** 2))
        blobs.append(sc.astype(np.float32))
    im = make_image(rng, H, W, image)
    tags = np.where(labels[1:] == 1)[0] + 1
    return dict(blobs=blobs, image=im, tags=tags)
