"""TEST INFRASTRUCTURE ONLY -- numpy restatement of AnnotationLayer.forward
(pylayers/pylayers/pylayers.py:369-387), the data layer that turns image tags and sparse localisation
cues into the dense blobs the hot path consumes.  Only tests/ may import it.

Parity status: pinned against the reference's own AnnotationLayer.forward executed in place
(oracle/ref_layers.py:annotation_layer_forward -> tests/golden/layers_ref.npz, compared bit for bit in
tests/test_oracle_golden.py); the statements below follow the reference's numpy statements in order,
including the order of the np.random draws.
"""
import numpy as np


def annotation_forward(data_file, image_ids, images, is_mirror, M=21, h=41, w=41):
    n = len(image_ids)
    top0 = np.zeros((n, 1, 1, M), np.float32)                       # :371
    top1 = np.zeros((n, M, h, w), np.float32)                       # :372
    top2 = np.array(images, np.float32)                             # :373
    for i, image_id in enumerate(image_ids):                        # :375
        labels_i = data_file['%i_labels' % image_id]                # :377
        top0[i, 0, 0, 0] = 1.0                                      # :378
        top0[i, 0, 0, labels_i] = 1.0                               # :379
        cues_i = data_file['%i_cues' % image_id]                    # :381
        top1[i, cues_i[0], cues_i[1], cues_i[2]] = 1.0              # :382
        if is_mirror:                                               # :384
            flip = np.random.choice(2) * 2 - 1                      # :385
            top1[i, ...] = top1[i, :, :, ::flip].copy()             # :386 (numpy buffers the overlap)
            top2[i, ...] = top2[i, :, :, ::flip].copy()             # :387
    return top0, top1, top2
