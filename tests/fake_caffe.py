"""tests' name for dsrg_b200/dropin/caffe_shim.py (the pycaffe stand-in the drop-in layers are driven through when
real Caffe is absent; bench.py's ``e2e_layer`` leg uses the same shim)."""
from dsrg_b200.dropin.caffe_shim import Blob, Layer, install, run_layer  # noqa: F401
