"""dsrg_b200/dropin/caffe_shim.py -- a minimal stand-in for pycaffe's Python-layer protocol (test pyramid item 3):
``caffe.Layer`` with ``param_str`` and Blob objects exposing ``data`` / ``diff`` / ``reshape`` --
exactly what external Caffe's PythonLayer shim hands to setup/reshape/forward/backward."""
import sys
import types

import numpy as np


class Blob(object):
    def __init__(self, data=None):
        self.data = np.zeros((1,), np.float32) if data is None else np.array(data, np.float32)
        self.diff = np.zeros_like(self.data)

    def reshape(self, *shape):
        if tuple(shape) != self.data.shape:
            self.data = np.zeros(shape, np.float32)
            self.diff = np.zeros(shape, np.float32)


class Layer(object):
    param_str = ""


def install():
    """Put a fake ``caffe`` module into sys.modules and the drop-in packages on sys.path (what
    training/tools/findcaffe.py:27-28 does for the reference's pylayers)."""
    import os
    if "caffe" not in sys.modules:
        m = types.ModuleType("caffe")
        m.Layer = Layer
        m.Blob = Blob
        sys.modules["caffe"] = m
    dropin = os.path.dirname(os.path.abspath(__file__))
    if dropin not in sys.path:
        sys.path.insert(0, dropin)


def run_layer(layer_cls, bottoms, param_str="", n_top=1):
    """setup -> reshape -> forward, the order Caffe uses.  Returns (layer, bottom blobs, top blobs)."""
    layer = layer_cls()
    layer.param_str = param_str
    bottom = [Blob(b) for b in bottoms]
    top = [Blob() for _ in range(n_top)]
    layer.setup(bottom, top)
    layer.reshape(bottom, top)
    layer.forward(bottom, top)
    return layer, bottom, top
