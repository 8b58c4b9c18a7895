"""GPU tests of the AnnotationLayer step: dsrg_annotation_forward_* and the drop-in
layer against oracle/annot_oracle.py (the reference's numpy statements)."""
import pickle

import numpy as np
import pytest

import fake_caffe

fake_caffe.install()
import pylayers  # noqa: E402
from dsrg_b200 import api  # noqa: E402
from oracle import annot_oracle  # noqa: E402

pytestmark = pytest.mark.gpu


def cue_file(n_images, seed=0, h=41, w=41, M=21):
    rng = np.random.RandomState(seed)
    d = {}
    for i in range(n_images):
        k = rng.randint(1, 4)
        tags = np.sort(rng.choice(np.arange(1, M), size=k, replace=False))
        d['%i_labels' % i] = tags
        cls = np.concatenate([[0], tags])
        K = 0 if i == 3 else rng.randint(1, 400)
        d['%i_cues' % i] = np.stack([rng.choice(cls, K), rng.randint(0, h, K), rng.randint(0, w, K)]).astype(np.int64)
    return d


@pytest.mark.parametrize("mirror", [False, True])
def test_annotation_layer_matches_reference_statements(torch_cuda, tmp_path, mirror):
    d = cue_file(8)
    with open(str(tmp_path / "cues.pickle"), "wb") as f:
        pickle.dump(d, f, protocol=2)
    ids = np.array([5, 0, 3, 7, 2, 2], np.float32).reshape(-1, 1, 1, 1)
    rng = np.random.RandomState(1)
    images = (rng.rand(6, 3, 33, 47) * 255 - 110).astype(np.float32)
    np.random.seed(11)
    want = annot_oracle.annotation_forward(d, ids.reshape(-1), images, mirror)
    np.random.seed(11)
    layer, bottom, top = fake_caffe.run_layer(
        pylayers.AnnotationLayer, [ids, images],
        param_str="{'cues': 'cues.pickle', 'mirror': %s, 'root': '%s'}" % (mirror, str(tmp_path)), n_top=3)
    for t, wnt in zip(top, want):
        assert t.data.shape == wnt.shape and np.array_equal(t.data, wnt)
    assert top[1].data.sum() > 0
    if mirror:   # the seeded draws flipped some images and not others
        flipped = [not np.array_equal(top[2].data[i], images[i]) for i in range(6)]
        assert any(flipped) and not all(flipped)


def test_annotation_dev_entry_negative_indices_and_errors(torch_cuda):
    import torch
    eng = api.Engine(4, 41, 41, 21)
    tags = [np.array([3, 7]), np.array([], np.int64), np.array([-1]), np.array([20])]
    cues = [np.array([[3, 7, 0], [0, 40, -1], [5, -41, 40]]), np.zeros((3, 0), np.int64),
            np.array([[-1], [-1], [-1]]), np.array([[20, 20], [1, 1], [2, 2]])]
    labels = torch.full((4, 1, 1, 21), 9.0, device="cuda")
    dense = torch.full((4, 21, 41, 41), 9.0, device="cuda")
    eng.annotation_forward_dev(tags, cues, labels, dense, flip=[0, 1, 1, 0])
    torch.cuda.synchronize()
    d = {}
    for i in range(4):
        d['%i_labels' % i], d['%i_cues' % i] = tags[i], cues[i]
    want0, want1, _ = annot_oracle.annotation_forward(d, range(4), np.zeros((4, 3, 1, 1), np.float32), False)
    want1[1] = want1[1][:, :, ::-1]
    want1[2] = want1[2][:, :, ::-1]
    assert np.array_equal(labels.cpu().numpy(), want0) and np.array_equal(dense.cpu().numpy(), want1)
    for bad in (np.array([[21], [0], [0]]), np.array([[0], [41], [0]]), np.array([[0], [0], [-42]])):
        with pytest.raises(api.DsrgError):
            eng.annotation_forward_host([np.array([1])], [bad])
    with pytest.raises(api.DsrgError):
        eng.annotation_forward_host([np.array([21])], [np.zeros((3, 0), np.int64)])
    eng.close()
