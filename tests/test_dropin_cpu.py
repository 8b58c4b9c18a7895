"""CPU tests of the drop-in boundary: the `pylayers` / `krahenbuhl2013` modules import under a fake
caffe, expose the reference's names and honour its setup/reshape/error contract.  No GPU compute."""
import inspect

import numpy as np
import pytest

import fake_caffe

fake_caffe.install()
import krahenbuhl2013  # noqa: E402
import pylayers  # noqa: E402


def test_module_surface_matches_reference_names():
    for name in ("SoftmaxLayer", "CRFLayer", "DSRGLayer", "BalancedSeedLossLayer", "ConstrainLossLayer",
                 "generate_seed_step", "min_prob"):
        assert hasattr(pylayers, name), name
    assert pylayers.min_prob == 0.0001
    sig = inspect.signature(krahenbuhl2013.CRF)
    assert list(sig.parameters) == ["image", "unary", "maxiter", "scale_factor", "color_factor"]
    assert [sig.parameters[k].default for k in ("maxiter", "scale_factor", "color_factor")] == [10, 1.0, 13]
    from krahenbuhl2013.wrapper import DenseCRF
    for m in ("set_unary_energy", "add_pairwise_energy", "map", "inference"):
        assert callable(getattr(DenseCRF, m))
    for cls in (pylayers.CRFLayer, pylayers.DSRGLayer, pylayers.BalancedSeedLossLayer):
        for m in ("setup", "reshape", "forward", "backward"):
            assert callable(getattr(cls, m))


@pytest.mark.parametrize("cls,need", [(pylayers.CRFLayer, 2), (pylayers.DSRGLayer, 4),
                                      (pylayers.BalancedSeedLossLayer, 2), (pylayers.ConstrainLossLayer, 2),
                                      (pylayers.SoftmaxLayer, 1)])
def test_wrong_bottom_count_raises_plain_exception(cls, need):
    layer = cls()
    layer.param_str = "{'th1': 0.99, 'th2': 0.85}"
    with pytest.raises(Exception):
        layer.setup([fake_caffe.Blob()] * (need + 1), [fake_caffe.Blob()])


def test_dsrg_layer_param_str_and_reshape():
    layer = pylayers.DSRGLayer()
    layer.param_str = "{'th1': 0.99, 'th2': 0.85}"           # train-s.prototxt:783
    bottom = [fake_caffe.Blob(np.zeros(s)) for s in ((2, 1, 1, 21), (2, 21, 41, 41), (2, 21, 41, 41), (2, 3, 321, 321))]
    top = [fake_caffe.Blob()]
    layer.setup(bottom, top)
    assert (layer._th1, layer._th2, layer._max_iters, layer._iter_index) == (0.99, 0.85, -1, 0)
    layer.reshape(bottom, top)
    assert top[0].data.shape == (2, 21, 41, 41)
    top[0].diff[...] = 3.0
    layer.backward(top, [True], bottom)                        # pylayers.py:307-308
    assert np.all(bottom[1].diff == 3.0)
    layer2 = pylayers.DSRGLayer()
    layer2.param_str = "{'th1': 0.9, 'th2': 0.8, 'iters': 7}"
    layer2.setup(bottom, top)
    assert layer2._max_iters == 7


def _num_grad(f, x, idx, eps=1e-3):
    xp, xm = x.copy(), x.copy()
    xp[idx] += eps
    xm[idx] -= eps
    return (f(xp) - f(xm)) / (2 * eps)


def test_softmax_and_constrain_loss_oracles_are_self_consistent():
    """The numpy restatements the GPU layers are checked against: analytic backward == numerical gradient."""
    from oracle import loss_oracle
    rng = np.random.RandomState(1)
    x = rng.randn(2, 5, 3, 4)
    g = rng.randn(*x.shape)
    p = loss_oracle.softmax_layer_forward(x)
    np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-12)
    gx = loss_oracle.softmax_layer_backward(x, g)
    for idx in [(0, 1, 2, 3), (1, 4, 0, 0)]:
        assert abs(_num_grad(lambda q: float(np.sum(loss_oracle.softmax_layer_forward(q) * g)), x, idx, 1e-5) - gx[idx]) < 1e-7
    pr = rng.rand(2, 4, 3, 3) * 0.8 + 0.1
    ls = np.log(rng.rand(2, 4, 3, 3) * 0.8 + 0.1)
    ls[0, 0, 0, 0] = np.log(1e-3)          # ratio below 0.05: clipped branch
    gp, gl = loss_oracle.constrain_loss_grad(pr, ls)
    for idx in [(0, 1, 2, 2), (1, 3, 0, 1), (0, 0, 0, 0)]:
        assert abs(_num_grad(lambda q: loss_oracle.constrain_loss(q, ls), pr, idx, 1e-6) - gp[idx]) < 1e-6
        assert abs(_num_grad(lambda q: loss_oracle.constrain_loss(pr, q), ls, idx, 1e-6) - gl[idx]) < 1e-6


def test_annotation_layer_and_postprocess_surface(tmp_path):
    """AnnotationLayer keeps the reference's setup / reshape contract (pylayers.py:348-367) without touching the
    GPU; the post-processing module exposes the two predict_mask() tails."""
    import pickle
    assert hasattr(pylayers, "AnnotationLayer")
    layer = pylayers.AnnotationLayer()
    layer.param_str = "{}"
    with pytest.raises(Exception):                       # "The layer needs two inputs!"
        layer.setup([fake_caffe.Blob()], [fake_caffe.Blob()] * 3)
    with open(str(tmp_path / "localization_cues.pickle"), "wb") as f:
        pickle.dump({"0_labels": np.array([3]), "0_cues": np.zeros((3, 0), np.int64)}, f, protocol=2)
    layer.param_str = "{'root': '%s'}" % str(tmp_path)   # default file name and mirror=False like the reference
    bottom = [fake_caffe.Blob(np.zeros((2, 1, 1, 1))), fake_caffe.Blob(np.zeros((2, 3, 33, 47)))]
    top = [fake_caffe.Blob(), fake_caffe.Blob(), fake_caffe.Blob()]
    layer.setup(bottom, top)
    assert layer._cue_name == "localization_cues.pickle" and layer.is_mirror is False
    layer.reshape(bottom, top)
    assert [t.data.shape for t in top] == [(2, 1, 1, 21), (2, 21, 41, 41), (2, 3, 33, 47)]
    from dsrg_b200 import postprocess
    assert list(inspect.signature(postprocess.predict_mask_ms).parameters)[:3] == ["im", "scores_per_scale", "smooth"]
    assert list(inspect.signature(postprocess.predict_mask_gt).parameters)[:4] == ["im", "scores", "labels", "smooth"]
    assert postprocess.EPS == 0.00001
