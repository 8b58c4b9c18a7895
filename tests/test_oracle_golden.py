"""CPU tests: the oracles against the committed golden vectors (generated from the reference's
own code by tests/golden/make_golden.py) and against each other."""
import numpy as np
import pytest

from conftest import load_golden
from dsrg_b200 import synth
from helpers import crf_case_inputs, digest, make_golden, srg_case_inputs
from oracle import crf_oracle, loss_oracle, srg_oracle

SRG_NAMES = [c[0] for c in make_golden.SRG_CASES]


@pytest.mark.parametrize("name", SRG_NAMES)
def test_srg_closed_form_matches_reference_golden(name):
    g = load_golden("srg_ref_%s.npz" % name)
    labels, cues, probs = srg_case_inputs(name)
    assert digest(labels, cues, probs) == str(g["inputs_sha256"]), "synthetic generator drifted"
    want = np.unpackbits(g["seeds_bits"])[: int(np.prod(g["seeds_shape"]))].reshape(g["seeds_shape"]).astype(np.float32)
    got = srg_oracle.srg_closed_form(labels, cues, probs, float(g["th1"]), float(g["th2"]))
    assert got.dtype == np.float32 and np.array_equal(got, want)


@pytest.mark.parametrize("name", [n for n in SRG_NAMES if n not in ("i", "j", "h")])
def test_srg_faithful_port_matches_reference_golden(name):
    g = load_golden("srg_ref_%s.npz" % name)
    labels, cues, probs = srg_case_inputs(name)
    want = np.unpackbits(g["seeds_bits"])[: int(np.prod(g["seeds_shape"]))].reshape(g["seeds_shape"]).astype(np.float32)
    got = srg_oracle.srg_faithful(labels, cues, probs, 0.99, 0.85)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not srg_oracle.reference_available(), reason="needs /root/reference (dev container)")
def test_srg_reference_in_place_random_sweep():
    from dsrg_b200 import synth
    for i in range(6):
        p = synth.make_problem(500 + i, 19 + 3 * i, 31 - 2 * i, cues=["cam", "random"][i % 2], image="noise")
        want = srg_oracle.run_reference(p["labels"], p["cues"], p["probs"], 0.99, 0.85)
        assert np.array_equal(srg_oracle.srg_closed_form(p["labels"], p["cues"], p["probs"], 0.99, 0.85), want)
        assert np.array_equal(srg_oracle.srg_faithful(p["labels"], p["cues"], p["probs"], 0.99, 0.85), want)


@pytest.mark.parametrize("case", make_golden.LATTICE_CASES, ids=[c[0] for c in make_golden.LATTICE_CASES])
def test_oracle_lattice_matches_reference_permutohedral_golden(case):
    """oracle/crf_oracle.c lattice vs the golden digests of the reference's own permutohedral.cpp."""
    from dsrg_b200 import synth
    name, H, W, var, sxy, srgb = case
    g = load_golden("lattice_ref.npz")
    im = None if var is None else synth.make_image(np.random.RandomState(77), H, W, var)
    f = make_golden.features(H, W, im, sxy, sxy, srgb)
    L = crf_oracle.OracleLattice(f)
    assert L.M == int(g[name + "/M"])
    assert digest(L.offset, L.bary, L.rank, L.n1, L.n2) == str(g[name + "/sha_struct"])
    x = np.random.RandomState(5).rand(H * W, 21).astype(np.float32)
    assert digest(L.compute(x, "sse")) == str(g[name + "/sha_sse"])
    assert digest(L.compute(np.ones((H * W, 1), np.float32), "seq")) == str(g[name + "/sha_seq"])
    np.testing.assert_allclose(L.bary.sum(1), 1.0, atol=2e-6)  # barycentric weights sum to one


@pytest.mark.skipif(not crf_oracle.ref_available(), reason="oracle/_ref not built")
def test_oracle_lattice_bit_exact_vs_ref_so_live():
    rng = np.random.RandomState(3)
    for H, W, d in ((23, 31, 2), (23, 31, 5), (40, 17, 5)):
        f = (rng.rand(H * W, d) * [3, 3, 19, 19, 19][:d]).astype(np.float32)
        a, b = crf_oracle.OracleLattice(f), crf_oracle.RefLattice(f)
        assert a.M == b.M
        for k in ("offset", "bary", "rank", "n1", "n2"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        x = rng.rand(H * W, 21).astype(np.float32)
        assert np.array_equal(a.compute(x, "sse"), b.compute(x, "dispatch"))


@pytest.mark.parametrize("name", [c[0] for c in make_golden.CRF_CASES])
def test_crf_oracle_matches_frozen_golden(name):
    g = load_golden("crf_oracle_%s.npz" % name)
    im, unary, sf = crf_case_inputs(name)
    assert digest(im, unary) == str(g["inputs_sha256"])
    q = crf_oracle.CRF(im, unary, maxiter=10, scale_factor=sf)
    np.testing.assert_allclose(q, g["Q"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(q.sum(-1), 1.0, atol=1e-5)


def test_crf_oracle_map_and_object_api():
    im, unary, sf = crf_case_inputs("train41_smooth")
    H, W, M = unary.shape
    c = crf_oracle.DenseCRF(W, H, M)
    c.set_unary_energy(-unary.ravel())
    c.add_pairwise_energy(10, 80 / sf, 80 / sf, 13, 13, 13, 3, 3 / sf, 3 / sf, im.ravel())
    q = c.inference(5).reshape(H, W, M)
    assert np.array_equal(c.map(5).reshape(H, W), q.argmax(-1))
    assert c.lattice(0).M > 0 and c.norm(1).shape == (H * W,)


def test_loss_oracle_gradient_is_consistent():
    rng = np.random.RandomState(0)
    p = rng.rand(2, 5, 6, 7) * 0.9 + 0.05
    lab = (rng.rand(2, 5, 6, 7) < 0.2).astype(np.float64)
    lab[1, 1:] = 0  # an image without foreground seeds: max(count, 1e-4) branch
    g = loss_oracle.balanced_seed_loss_grad(p, lab)
    eps = 1e-6
    for idx in [(0, 0, 1, 2), (0, 3, 2, 2), (1, 0, 5, 6)]:
        pp, pm = p.copy(), p.copy()
        pp[idx] += eps
        pm[idx] -= eps
        num = (loss_oracle.balanced_seed_loss(pp, lab) - loss_oracle.balanced_seed_loss(pm, lab)) / (2 * eps)
        assert abs(num - g[idx]) < 1e-5


def test_zoom_restatement_is_bit_exact_vs_scipy():
    """The bilinear-zoom restatement the CUDA preprocessing kernel mirrors, against scipy itself."""
    from scipy.ndimage import zoom
    rng = np.random.RandomState(4)
    for (ih, iw, oh, ow) in ((321, 321, 41, 41), (97, 131, 33, 57), (50, 40, 50, 40), (17, 9, 40, 31), (8, 8, 1, 1),
                             (5, 7, 1, 4), (33, 33, 321, 321), (31, 31, 97, 131), (51, 51, 60, 45)):
        im = (rng.rand(2, 3, ih, iw) * 255 - 110).astype(np.float32)
        want = zoom(im, (1.0, 1.0, float(oh) / ih, float(ow) / iw), order=1)
        got = crf_oracle.zoom_order1_restated(im, oh, ow)
        assert want.shape == got.shape and want.dtype == got.dtype and np.array_equal(want, got)


def test_zoom_restatement_covers_score_maps():
    """(h,w,M) score maps zoomed with a trailing identity axis (test-ms.py:96): same taps, same bits."""
    from scipy.ndimage import zoom
    rng = np.random.RandomState(5)
    for (h, w, H, W) in ((41, 41, 375, 500), (31, 31, 97, 131), (51, 51, 60, 45), (41, 41, 20, 30), (7, 9, 1, 5),
                         (33, 33, 321, 321)):
        blob = (rng.randn(21, h, w) * 5).astype(np.float32)
        want = zoom(np.transpose(blob, [1, 2, 0]), (float(H) / h, float(W) / w, 1.0), order=1)
        got = crf_oracle.zoom_order1_restated(blob[None], H, W)[0].transpose(1, 2, 0)
        assert want.shape == (H, W, 21) and np.array_equal(want, got)


@pytest.mark.parametrize("case", make_golden.POST_CASES, ids=[c[0] for c in make_golden.POST_CASES])
def test_post_oracle_matches_frozen_golden(case):
    from oracle import post_oracle
    name, mode, H, W, sizes, index = case
    g = load_golden("post_oracle_%s.npz" % name)
    im, blobs, tags = make_golden.post_inputs(H, W, sizes, index)
    assert digest(im, *blobs) == str(g["inputs_sha256"])
    if mode == "ms":
        lab, probs = post_oracle.predict_mask_ms(im, blobs, smooth=True)
        sel = list(range(21))
    else:
        lab, probs = post_oracle.predict_mask_gt(im, blobs[0], tags, smooth=True)
        sel = [0] + list(tags)
        assert set(np.unique(lab)) <= set(sel)
    np.testing.assert_allclose(probs, g["probs"], atol=1e-6, rtol=0)
    bad = lab != g["labels"]
    if bad.any():   # another libm may flip an exact near-tie
        top2 = np.sort(g["probs"][:, :, sel], axis=2)[:, :, -2:]
        assert ((top2[:, :, 1] - top2[:, :, 0])[bad] < 1e-5).all()
    # smooth=False is plain numpy: arg-max of the clamped soft-max
    lab0, p0 = post_oracle.predict_mask_ms(im, blobs, smooth=False)
    assert np.array_equal(lab0, p0.argmax(2)) and p0.min() >= np.float32(1e-5)


needs_ref_crf = pytest.mark.skipif(not crf_oracle.ref_crf_available(),
                                   reason="oracle/_ref/libdensecrf_ref.so not built (needs /root/reference once)")


@needs_ref_crf
@pytest.mark.parametrize("name", [c[0] for c in make_golden.CRF_CASES])
def test_crf_oracle_is_bit_exact_vs_the_reference_build(name):
    """The C restatement against the reference's OWN CRF sources (densecrf.cpp, pairwise.cpp, ... compiled
    unmodified, oracle/Makefile) through DenseCRFWrapper: marginals and MAP labels, several iteration counts."""
    im, unary, sf = crf_case_inputs(name)
    H, W, M = unary.shape
    assert np.array_equal(crf_oracle.CRF(im, unary, 10, sf), crf_oracle.CRF_reference(im, unary, 10, sf))
    o, r = crf_oracle.DenseCRF(W, H, M), crf_oracle.RefDenseCRF(W, H, M)
    for c in (o, r):
        c.set_unary_energy(-unary.ravel())
        c.add_pairwise_energy(10, 80 / sf, 80 / sf, 13, 13, 13, 3, 3 / sf, 3 / sf, im.ravel())
    for it in (0, 1, 4):
        assert np.array_equal(o.inference(it), r.inference(it))
        assert np.array_equal(o.map(it), r.map(it))


@needs_ref_crf
@pytest.mark.parametrize("H,W,M,sf,img", [(97, 131, 21, 1.0, "smooth"), (60, 45, 5, 12.0, "noise"), (33, 33, 2, 1.0, "noise"),
                                          (161, 161, 21, 1.0, "smooth"), (120, 90, 21, 1.0, "photo")])
def test_crf_oracle_vs_reference_build_other_shapes(H, W, M, sf, img):
    """incl. M = 2, where Permutohedral::compute takes the seqCompute branch (permutohedral.cpp:600-601)."""
    rng = np.random.RandomState(H * 7 + M)
    im = synth.make_image(rng, H, W, img)
    logits = rng.randn(H, W, M) * 2
    logits[H // 4: H // 2, W // 3:, 0] += 4
    e = np.exp(logits - logits.max(2, keepdims=True))
    pr = (e / e.sum(2, keepdims=True)).astype(np.float32)
    pr[pr < 1e-5] = 1e-5
    for unary in (np.log(pr), pr):
        a = crf_oracle.CRF(im, unary, 10, sf, color_factor=13)
        b = crf_oracle.CRF_reference(im, unary, 10, sf, color_factor=13)
        assert np.array_equal(a, b)


# ---- a12 / f1 / f4: the loss and annotation oracles against the reference's own layer classes -------------------
# tests/golden/layers_ref.npz holds the outputs of the reference's SoftmaxLayer / BalancedSeedLossLayer /
# ConstrainLossLayer / AnnotationLayer class bodies executed in place (oracle/ref_layers.py; Theano is replaced by
# oracle/theano_shim.py, float64 and float32 evaluation).
@pytest.mark.parametrize("seed", [0, 1])
def test_loss_oracles_match_the_reference_layers_golden(seed):
    g = load_golden("layers_ref.npz")
    logits, p, lab, td, ls = make_golden.layer_inputs(seed)
    assert digest(logits, p, lab, td, ls) == str(g["in%d_sha256" % seed]), "input generator drifted"
    k = "s%d_float64_" % seed
    tight = dict(rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(loss_oracle.softmax_layer_forward(logits), g[k + "softmax_probs"], **tight)
    np.testing.assert_allclose(loss_oracle.softmax_layer_backward(logits, td), g[k + "softmax_grad"], **tight)
    np.testing.assert_allclose(loss_oracle.balanced_seed_loss(p, lab), g[k + "seed_loss"], **tight)
    np.testing.assert_allclose(loss_oracle.balanced_seed_loss_grad(p, lab), g[k + "seed_grad"], **tight)
    np.testing.assert_allclose(loss_oracle.constrain_loss(p, ls), g[k + "constrain_loss"], **tight)
    w0, w1 = loss_oracle.constrain_loss_grad(p, ls)
    np.testing.assert_allclose(w0, g[k + "constrain_g0"], **tight)
    np.testing.assert_allclose(w1, g[k + "constrain_g1"], **tight)
    # the float32 evaluation (what T.ftensor4 computes in) stays within the tolerance the GPU tests use
    k32 = "s%d_float32_" % seed
    for name in ("softmax_probs", "softmax_grad", "seed_loss", "seed_grad", "constrain_loss", "constrain_g0", "constrain_g1"):
        np.testing.assert_allclose(g[k32 + name], g[k + name], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("mirror", [False, True])
def test_annotation_oracle_matches_the_reference_layer_golden(mirror):
    from oracle import annot_oracle
    g = load_golden("layers_ref.npz")
    d, ids, images = make_golden.annot_inputs()
    np.random.seed(11)
    t0, t1, t2 = annot_oracle.annotation_forward(d, ids, images, mirror)
    k = "annot_m%d_" % int(mirror)
    shape = tuple(g[k + "cues_shape"])
    want1 = np.unpackbits(g[k + "cues_bits"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)
    assert np.array_equal(t0, g[k + "labels"]) and np.array_equal(t1, want1) and digest(t2) == str(g[k + "images_sha256"])


def _ref_layers():
    from oracle import ref_layers
    return ref_layers


@pytest.mark.skipif(not srg_oracle.reference_available(), reason="needs /root/reference (dev container)")
def test_reference_layers_run_in_place_random_sweep():
    """Live: the reference's own layer classes (through the Theano stand-in) against the restatements on fresh inputs."""
    rl = _ref_layers()
    for seed in (5, 6, 7):
        logits, p, lab, td, ls = make_golden.layer_inputs(seed, N=2 + seed % 3, H=5 + seed, W=11 - seed)
        loss, grad = rl.balanced_seed_loss_layer(p, lab, "float64")
        np.testing.assert_allclose(loss, loss_oracle.balanced_seed_loss(p, lab), rtol=1e-12)
        np.testing.assert_allclose(grad, loss_oracle.balanced_seed_loss_grad(p, lab), rtol=1e-12, atol=1e-15)
        pr, gd = rl.softmax_layer(logits, td, "float64")
        np.testing.assert_allclose(pr, loss_oracle.softmax_layer_forward(logits), rtol=1e-12)
        np.testing.assert_allclose(gd, loss_oracle.softmax_layer_backward(logits, td), rtol=1e-10, atol=1e-15)
        cl, g0, g1 = rl.constrain_loss_layer(p, ls, "float64")
        np.testing.assert_allclose(cl, loss_oracle.constrain_loss(p, ls), rtol=1e-12)
        w0, w1 = loss_oracle.constrain_loss_grad(p, ls)
        np.testing.assert_allclose(g0, w0, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(g1, w1, rtol=1e-12, atol=1e-15)
