"""oracle/crf_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the CPU checkers built by oracle/Makefile:

* ``liboracle_crf.so``  -- the plain-C restatement in oracle/crf_oracle.c
* ``_ref/libpermuto_ref.so`` -- the reference's own CRF/src/permutohedral.cpp (when built)
* ``_ref/libdensecrf_ref.so`` -- the reference's own CRF sources behind DenseCRFWrapper (when built):
  ``RefDenseCRF`` / ``CRF_reference``; the restatement is bit-identical to it (tests/test_oracle_golden.py)

and a Python restatement of the two reference call sites of the CRF:
``CRF()`` (CRF/krahenbuhl2013/CRF.py:4-37) and ``DSRGLayer.refinement`` /
``CRFLayer.forward`` (pylayers/pylayers/pylayers.py:310-331, :63-88).

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the checkers (gcc only; no GPU needed)."""
    so = os.path.join(_HERE, "liboracle_crf.so")
    src = os.path.join(_HERE, "crf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_crf.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/CRF/src/permutohedral.cpp"):
        refs = [os.path.join(_HERE, "_ref", n) for n in ("libpermuto_ref.so", "libdensecrf_ref.so")]
        if force or not all(os.path.exists(r) for r in refs):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle_crf.so"))
        L.oracle_lattice_init.restype = C.c_void_p
        L.oracle_lattice_init.argtypes = [_f32p, C.c_int, C.c_int]
        L.oracle_lattice_free.argtypes = [C.c_void_p]
        L.oracle_lattice_M.argtypes = [C.c_void_p]
        for n, t in (("offset", C.c_int), ("rank", C.c_int), ("n1", C.c_int), ("n2", C.c_int),
                     ("bary", C.c_float)):
            fn = getattr(L, "oracle_lattice_" + n)
            fn.restype = C.POINTER(t)
            fn.argtypes = [C.c_void_p]
        for n in ("oracle_lattice_seq_compute", "oracle_lattice_sse_compute"):
            getattr(L, n).argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
        L.oracle_crf_create.restype = C.c_void_p
        L.oracle_crf_create.argtypes = [C.c_int] * 3
        L.oracle_crf_destroy.argtypes = [C.c_void_p]
        L.oracle_crf_set_unary_energy.argtypes = [C.c_void_p, _f32p]
        L.oracle_crf_add_pairwise_energy.argtypes = [C.c_void_p] + [C.c_float] * 9 + [_u8p]
        L.oracle_crf_inference.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.oracle_crf_map.argtypes = [C.c_void_p, C.c_int, _i32p]
        L.oracle_crf_lattice.restype = C.c_void_p
        L.oracle_crf_lattice.argtypes = [C.c_void_p, C.c_int]
        L.oracle_crf_norm.restype = C.POINTER(C.c_float)
        L.oracle_crf_norm.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpermuto_ref.so"))


def ref():
    """The reference's own permutohedral.cpp (oracle/_ref), or None if it was never built."""
    global _REF
    if _REF is None and ref_available():
        R = C.CDLL(os.path.join(_HERE, "_ref", "libpermuto_ref.so"))
        R.ref_lattice_init.restype = C.c_void_p
        R.ref_lattice_init.argtypes = [_f32p, C.c_int, C.c_int]
        R.ref_lattice_free.argtypes = [C.c_void_p]
        R.ref_lattice_M.argtypes = [C.c_void_p]
        for n, t in (("offset", C.c_int), ("rank", C.c_int), ("bary", C.c_float)):
            fn = getattr(R, "ref_lattice_" + n)
            fn.restype = C.POINTER(t)
            fn.argtypes = [C.c_void_p]
        R.ref_lattice_neighbors.argtypes = [C.c_void_p, _i32p, _i32p]
        for n in ("ref_lattice_seq_compute", "ref_lattice_sse_compute"):
            getattr(R, n).argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
        R.ref_lattice_compute.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
        _REF = R
    return _REF


class _LatticeView(object):
    """Common read-out of a lattice: offsets/bary/rank [(N)(d+1)], neighbours [(d+1)][M]."""

    def _grab(self, ptr, n, dtype):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class OracleLattice(_LatticeView):
    """oracle_lattice_init: restatement of Permutohedral::init (permutohedral.cpp:140-321)."""

    def __init__(self, feature, handle=None, owner=True):
        L = lib()
        self._L = L
        self.owner = owner
        if handle is None:
            feature = np.ascontiguousarray(feature, np.float32)  # [N][d] == d x N col-major
            self.N, self.d = feature.shape
            self.h = L.oracle_lattice_init(feature, self.d, self.N)
        else:
            self.h = handle
            self.N, self.d = feature
        self.M = L.oracle_lattice_M(self.h)
        n = self.N * (self.d + 1)
        self.offset = self._grab(L.oracle_lattice_offset(self.h), n, np.int32).reshape(self.N, self.d + 1)
        self.bary = self._grab(L.oracle_lattice_bary(self.h), n, np.float32).reshape(self.N, self.d + 1)
        self.rank = self._grab(L.oracle_lattice_rank(self.h), n, np.int32).reshape(self.N, self.d + 1)
        m = self.M * (self.d + 1)
        self.n1 = self._grab(L.oracle_lattice_n1(self.h), m, np.int32).reshape(self.d + 1, self.M)
        self.n2 = self._grab(L.oracle_lattice_n2(self.h), m, np.int32).reshape(self.d + 1, self.M)

    def compute(self, x, kind="sse"):
        x = np.ascontiguousarray(x, np.float32)  # [N][vs]
        out = np.empty_like(x)
        fn = self._L.oracle_lattice_sse_compute if kind == "sse" else self._L.oracle_lattice_seq_compute
        fn(self.h, out, x, x.shape[1])
        return out

    def __del__(self):
        if getattr(self, "owner", False) and getattr(self, "h", None):
            self._L.oracle_lattice_free(self.h)
            self.h = None


class RefLattice(_LatticeView):
    """The real reference lattice (oracle/_ref)."""

    def __init__(self, feature):
        R = ref()
        if R is None:
            raise RuntimeError("oracle/_ref/libpermuto_ref.so was not built")
        self._R = R
        feature = np.ascontiguousarray(feature, np.float32)
        self.N, self.d = feature.shape
        self.h = R.ref_lattice_init(feature, self.d, self.N)
        self.M = R.ref_lattice_M(self.h)
        n = self.N * (self.d + 1)
        self.offset = self._grab(R.ref_lattice_offset(self.h), n, np.int32).reshape(self.N, self.d + 1)
        self.bary = self._grab(R.ref_lattice_bary(self.h), n, np.float32).reshape(self.N, self.d + 1)
        self.rank = self._grab(R.ref_lattice_rank(self.h), n, np.int32).reshape(self.N, self.d + 1)
        self.n1 = np.empty((self.d + 1, self.M), np.int32)
        self.n2 = np.empty((self.d + 1, self.M), np.int32)
        R.ref_lattice_neighbors(self.h, self.n1, self.n2)

    def compute(self, x, kind="sse"):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        if kind == "dispatch":
            self._R.ref_lattice_compute(self.h, out, x, x.shape[1], self.N)
        else:
            fn = self._R.ref_lattice_sse_compute if kind == "sse" else self._R.ref_lattice_seq_compute
            fn(self.h, out, x, x.shape[1])
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self._R.ref_lattice_free(self.h)
            self.h = None


class DenseCRF(object):
    """Same surface as the reference's Cython type (CRF/krahenbuhl2013/wrapper.pyx:20-60)."""

    def __init__(self, W, H, nlabels):
        self._L = lib()
        self.W, self.H, self.M = int(W), int(H), int(nlabels)
        self.h = self._L.oracle_crf_create(self.W, self.H, self.M)

    def set_unary_energy(self, unary_costs):
        u = np.ascontiguousarray(unary_costs, np.float32)
        assert u.size == self.W * self.H * self.M
        self._L.oracle_crf_set_unary_energy(self.h, u)

    def add_pairwise_energy(self, w1, ta1, ta2, tb1, tb2, tb3, w2, tg1, tg2, im):
        im = np.ascontiguousarray(im, np.uint8)
        assert im.size == self.W * self.H * 3
        self._L.oracle_crf_add_pairwise_energy(self.h, w1, ta1, ta2, tb1, tb2, tb3, w2, tg1, tg2, im)

    def inference(self, n_iters=10):
        out = np.empty(self.W * self.H * self.M, np.float32)
        self._L.oracle_crf_inference(self.h, int(n_iters), out)
        return out

    def map(self, n_iters=10):
        out = np.empty(self.W * self.H, np.int32)
        self._L.oracle_crf_map(self.h, int(n_iters), out)
        return out

    def lattice(self, k):
        """k=0 spatial (gaussian), k=1 bilateral -- the order densecrf_wrapper.cpp:25-29 adds them."""
        d = 2 if k == 0 else 5
        return OracleLattice((self.W * self.H, d), handle=self._L.oracle_crf_lattice(self.h, k), owner=False)

    def norm(self, k):
        return np.ctypeslib.as_array(self._L.oracle_crf_norm(self.h, k), shape=(self.W * self.H,)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            self._L.oracle_crf_destroy(self.h)
            self.h = None


_REFCRF = None


def ref_crf_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdensecrf_ref.so"))


def _ref_crf():
    global _REFCRF
    if _REFCRF is None:
        R = C.CDLL(os.path.join(_HERE, "_ref", "libdensecrf_ref.so"))
        R.ref_densecrf_create.restype = C.c_void_p
        R.ref_densecrf_create.argtypes = [C.c_int] * 3
        R.ref_densecrf_destroy.argtypes = [C.c_void_p]
        R.ref_densecrf_set_unary_energy.argtypes = [C.c_void_p, _f32p]
        R.ref_densecrf_add_pairwise_energy.argtypes = [C.c_void_p] + [C.c_float] * 9 + [_u8p]
        R.ref_densecrf_inference.argtypes = [C.c_void_p, C.c_int, _f32p]
        R.ref_densecrf_map.argtypes = [C.c_void_p, C.c_int, _i32p]
        _REFCRF = R
    return _REFCRF


class RefDenseCRF(object):
    """The reference's OWN DenseCRFWrapper (CRF/src/*.cpp compiled in place into oracle/_ref/libdensecrf_ref.so
    against the Eigen stand-in); same surface as wrapper.pyx:20-60.  Present wherever oracle/_ref was built."""

    def __init__(self, W, H, nlabels):
        self._R = _ref_crf()
        self.W, self.H, self.M = int(W), int(H), int(nlabels)
        self.h = self._R.ref_densecrf_create(self.W, self.H, self.M)

    def set_unary_energy(self, unary_costs):
        u = np.ascontiguousarray(unary_costs, np.float32)
        assert u.size == self.W * self.H * self.M
        self._R.ref_densecrf_set_unary_energy(self.h, u)

    def add_pairwise_energy(self, w1, ta1, ta2, tb1, tb2, tb3, w2, tg1, tg2, im):
        im = np.ascontiguousarray(im, np.uint8)
        assert im.size == self.W * self.H * 3
        self._R.ref_densecrf_add_pairwise_energy(self.h, w1, ta1, ta2, tb1, tb2, tb3, w2, tg1, tg2, im)

    def inference(self, n_iters=10):
        out = np.empty(self.W * self.H * self.M, np.float32)
        self._R.ref_densecrf_inference(self.h, int(n_iters), out)
        return out

    def map(self, n_iters=10):
        out = np.empty(self.W * self.H, np.int32)
        self._R.ref_densecrf_map(self.h, int(n_iters), out)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self._R.ref_densecrf_destroy(self.h)
            self.h = None


def CRF_reference(image, unary, maxiter=10, scale_factor=1.0, color_factor=13):
    """CRF/krahenbuhl2013/CRF.py:4-37 on top of the reference's own compiled DenseCRFWrapper."""
    assert image.shape[:2] == unary.shape[:2]
    H, W = image.shape[:2]
    nlabels = unary.shape[2]
    crf = RefDenseCRF(W, H, nlabels)
    crf.set_unary_energy(-unary.ravel().astype("float32"))
    crf.add_pairwise_energy(10, 80 / scale_factor, 80 / scale_factor, color_factor, color_factor, color_factor,
                            3, 3 / scale_factor, 3 / scale_factor, image.ravel().astype("ubyte"))
    return crf.inference(maxiter).reshape((H, W, nlabels))


def CRF(image, unary, maxiter=10, scale_factor=1.0, color_factor=13):
    """Restatement of CRF/krahenbuhl2013/CRF.py:4-37 on top of the oracle DenseCRF."""
    assert image.shape[:2] == unary.shape[:2]
    H, W = image.shape[:2]
    nlabels = unary.shape[2]
    crf = DenseCRF(W, H, nlabels)
    crf.set_unary_energy(-unary.ravel().astype("float32"))
    crf.add_pairwise_energy(10, 80 / scale_factor, 80 / scale_factor, color_factor, color_factor, color_factor,
                            3, 3 / scale_factor, 3 / scale_factor, image.ravel().astype("ubyte"))
    return crf.inference(maxiter).reshape((H, W, nlabels))


MIN_PROB = 0.0001  # pylayers/pylayers/pylayers.py:20


def prepare_image(im, h, w):
    """Image preprocessing of pylayers.py:70-75 / :315-319 (zoom order=1, +mean, np.round)."""
    from scipy.ndimage import zoom
    mean_pixel = np.array([104.0, 117.0, 123.0])
    im = zoom(im, (1.0, 1.0, float(h) / im.shape[2], float(w) / im.shape[3]), order=1)
    im = np.transpose(im, [0, 2, 3, 1])
    im = im + mean_pixel[None, None, None, :]
    return np.round(im)


def zoom_order1_restated(im, oh, ow):
    """scipy.ndimage.zoom(im, (1, 1, oh/H, ow/W), order=1) for a (N,C,H,W) float32 array, restated
    operation by operation; tests check it bit-for-bit against scipy and the CUDA kernels against it.
    scipy's arithmetic (ni_interpolation.c NI_ZoomShift, ni_splines.c), all float64:
      * output index o maps to c = o * ((in-1)/(out-1)); c > in-1 (possible by one rounding error at the
        last row / column) is out of bounds under the default mode='constant' -> the pixel is cval = 0
      * weights w0 = 1 - (c - floor(c)), w1 = 1 - w0   (NOT c - floor(c))
      * each tap is ((v * wy) * wx), taps summed in the order (y0,x0), (y0,x1), (y1,x0), (y1,x1)
      * the sum is cast to the input dtype
    """
    n, c, ih, iw = im.shape
    zy = (ih - 1) / (oh - 1) if oh > 1 else 0.0
    zx = (iw - 1) / (ow - 1) if ow > 1 else 0.0
    ys = np.arange(oh, dtype=np.float64) * zy
    xs = np.arange(ow, dtype=np.float64) * zx
    oob = (ys[:, None] > ih - 1) | (xs[None, :] > iw - 1)
    y0 = np.minimum(np.floor(ys).astype(int), ih - 1)
    x0 = np.minimum(np.floor(xs).astype(int), iw - 1)
    wy0, wx0 = 1 - (ys - y0), 1 - (xs - x0)
    wy1, wx1 = 1 - wy0, 1 - wx0
    y1, x1 = np.minimum(y0 + 1, ih - 1), np.minimum(x0 + 1, iw - 1)
    a = im.astype(np.float64)

    def tap(yy, xx, wy, wx):
        return (a[:, :, yy][:, :, :, xx] * wy[:, None]) * wx[None, :]

    t = tap(y0, x0, wy0, wx0)
    t = t + tap(y0, x1, wy0, wx1)
    t = t + tap(y1, x0, wy1, wx0)
    t = t + tap(y1, x1, wy1, wx1)
    t[:, :, oob] = 0.0
    return t.astype(im.dtype)


def renormalise(q_nchw):
    """pylayers.py:328-330 on raw float32 CRF marginals given as (N,C,H,W) or (C,H,W): the clamp and the
    float64 renormalisation IN THE REFERENCE'S MEMORY LAYOUT -- ``result`` is an (N,H,W,C) float64 array seen
    through a transposed view, so np.sum(axis=1) reduces the contiguous class axis with NumPy's pairwise order
    (8 accumulators + tail), which differs from a sum over the outer axis of a planar array in the last bit."""
    q = np.asarray(q_nchw)
    single = q.ndim == 3
    if single:
        q = q[None]
    result = np.zeros((q.shape[0], q.shape[2], q.shape[3], q.shape[1]))        # :323 (float64, N,H,W,C)
    result[...] = np.transpose(q, (0, 2, 3, 1))                                  # :326 (float32 CRF output)
    result = np.transpose(result, [0, 3, 1, 2])                                  # :328
    result[result < MIN_PROB] = MIN_PROB                                         # :329
    result = result / np.sum(result, axis=1, keepdims=True)                      # :330
    return result[0] if single else result


def refinement(probs, im, scale_factor=12.0):
    """DSRGLayer.refinement, pylayers.py:310-331.  Mutates ``probs`` in place like the reference
    (:312).  Returns the float64 N x C x h x w renormalised CRF marginals."""
    _, _, h, w = probs.shape
    probs[probs < MIN_PROB] = MIN_PROB
    unary = np.transpose(np.array(probs), [0, 2, 3, 1])
    im = prepare_image(im, h, w)
    N = unary.shape[0]
    result = np.zeros(unary.shape)
    for i in range(N):
        result[i] = CRF(im[i], unary[i], scale_factor=scale_factor)
    result = np.transpose(result, [0, 3, 1, 2])
    result[result < MIN_PROB] = MIN_PROB
    result = result / np.sum(result, axis=1, keepdims=True)
    return result
