"""Python host side of the C ABI: thin wrappers, no compute.

`Engine` mirrors dsrg_engine.  Device arrays are torch CUDA tensors used purely as memory
containers (``data_ptr()``); host arrays are numpy.  Every method maps 1:1 onto an entry point of
include/dsrg_b200.h -- see that header for the reference interface each one replaces.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (LAYOUT_NCHW, LAYOUT_NHWC, POST_SUM_SCORES, POST_ZOOM_PROBS, CrfParams, DsrgError,  # noqa: F401
                   check)


def crf_params(scale_factor=1.0, color_factor=13, maxiter=10):
    """The pairwise parameters CRF() uses (CRF/krahenbuhl2013/CRF.py:31-32)."""
    p = CrfParams()
    _lib.lib().dsrg_crf_params_default(C.byref(p), float(scale_factor), float(color_factor), int(maxiter))
    return p


def _dptr(t):
    """Device pointer of a contiguous torch CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _hptr(a, dtype):
    if a is None:
        return None
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("expected a C-contiguous numpy array of dtype %s" % np.dtype(dtype))
    return C.c_void_p(a.ctypes.data)


def _stream(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(int(stream))


class _PinnedBlock(object):
    """Owner of one dsrg_host_alloc block; the numpy arrays built over it keep it alive through the ctypes
    buffer they wrap (numpy's ``base`` chain), and the block is returned with dsrg_host_free when the last of
    them is gone."""

    def __init__(self, L, nbytes):
        self._L = L
        self.ptr = L.dsrg_host_alloc(nbytes)
        if not self.ptr:
            raise DsrgError(_lib.E_NOMEM, L.dsrg_last_error().decode())

    def __del__(self):
        if getattr(self, "ptr", None):
            try:
                self._L.dsrg_host_free(self.ptr)
            except Exception:
                pass
            self.ptr = None


def pinned_empty(shape, dtype):
    """numpy array over pinned host memory from dsrg_host_alloc (freed when the array and its views are)."""
    L = _lib.lib()
    dtype = np.dtype(dtype)
    n = max(int(np.prod(shape)) * dtype.itemsize, 16)
    block = _PinnedBlock(L, n)
    buf = (C.c_char * n).from_address(block.ptr)
    buf._dsrg_block = block   # the ctypes object is the base of every array below: it carries the owner
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Engine(object):
    """dsrg_engine: all device buffers for batches of up to ``max_batch`` H x W x M problems."""

    def __init__(self, max_batch, H, W, M=21, device=None):
        """device None: the calling thread's current CUDA device (dsrg_current_device: what caffe.set_device /
        torch.cuda.set_device selected, or DSRG_B200_DEVICE)."""
        self._L = _lib.lib()
        if device is None:
            device = self._L.dsrg_current_device()
        self.max_batch, self.H, self.W, self.M, self.device = int(max_batch), int(H), int(W), int(M), int(device)
        self.h = self._L.dsrg_engine_create(self.device, self.max_batch, self.H, self.W, self.M)
        if not self.h:
            raise DsrgError(_lib.E_CUDA, self._L.dsrg_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self._L.dsrg_engine_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def device_bytes(self):
        return self._L.dsrg_engine_device_bytes(self.h)

    def set_size(self, H, W):
        """Select an image size within the capacity the engine was created with (no reallocation)."""
        check(self._L.dsrg_engine_set_size(self.h, int(H), int(W)))
        self.H, self.W = int(H), int(W)

    @property
    def capacity(self):
        v = [C.c_int() for _ in range(4)]
        check(self._L.dsrg_engine_get_size(self.h, *[C.byref(x) for x in v]))
        return v[2].value, v[3].value

    def set_host_chunk(self, images):
        check(self._L.dsrg_engine_set_host_chunk(self.h, int(images)))

    def set_graphs(self, enable):
        """CUDA-graph replay of repeated device passes (default on; needs a non-default stream)."""
        check(self._L.dsrg_engine_set_graphs(self.h, int(bool(enable))))

    @property
    def graph_replays(self):
        return int(self._L.dsrg_engine_graph_replays(self.h))

    def set_lanes(self, lanes):
        check(self._L.dsrg_engine_set_lanes(self.h, int(lanes)))

    @property
    def hybrid_tiles(self):
        """Tiles of the last mean-field pass that took the hybrid path (textured images); synchronises."""
        n = int(self._L.dsrg_engine_hybrid_tiles(self.h))
        if n < 0:
            raise DsrgError(_lib.last_error())
        return n

    def take_launch_count(self):
        return int(self._L.dsrg_engine_take_launch_count(self.h))

    # ---- device entry points (torch tensors as containers) ----
    def crf_dev(self, unary, image, params, out, unary_layout=LAYOUT_NHWC, out_layout=LAYOUT_NHWC, stream=None):
        B = image.shape[0]
        check(self._L.dsrg_crf_batch_dev(self.h, B, _dptr(unary), unary_layout, _dptr(image), C.byref(params),
                                         _dptr(out), out_layout, _stream(stream)))
        return out

    def crf_map_dev(self, unary, image, params, labels_out, unary_layout=LAYOUT_NHWC, stream=None):
        B = image.shape[0]
        check(self._L.dsrg_crf_map_batch_dev(self.h, B, _dptr(unary), unary_layout, _dptr(image), C.byref(params),
                                             _dptr(labels_out), _stream(stream)))
        return labels_out

    def srg_dev(self, labels, probs, cues, th1, th2, seeds_out, renorm=False, label_map_out=None, stream=None):
        B = probs.shape[0]
        check(self._L.dsrg_srg_batch_dev(self.h, B, _dptr(labels), _dptr(probs), _dptr(cues), float(th1), float(th2),
                                         int(bool(renorm)), _dptr(seeds_out), _dptr(label_map_out), _stream(stream)))
        return seeds_out

    def dsrg_forward_dev(self, labels, probs, cues, image, params, th1, th2, seeds_out, crf_out=None, stream=None):
        B = probs.shape[0]
        check(self._L.dsrg_dsrg_forward_dev(self.h, B, _dptr(labels), _dptr(probs), _dptr(cues), _dptr(image),
                                            C.byref(params), float(th1), float(th2), _dptr(seeds_out),
                                            _dptr(crf_out), _stream(stream)))
        return seeds_out

    def crflayer_forward_dev(self, probs, image, params, log_out, result=None, stream=None):
        B = probs.shape[0]
        check(self._L.dsrg_crflayer_forward_dev(self.h, B, _dptr(probs), _dptr(image), C.byref(params),
                                                _dptr(log_out), _dptr(result), _stream(stream)))
        return log_out

    def seedloss_forward_dev(self, probs, seeds, terms_out, stream=None):
        B = probs.shape[0]
        check(self._L.dsrg_seedloss_forward_dev(self.h, B, _dptr(probs), _dptr(seeds), _dptr(terms_out), _stream(stream)))
        return terms_out

    def seedloss_backward_dev(self, probs, seeds, grad_out, n_global=None, top_diff=1.0, stream=None):
        B = probs.shape[0]
        check(self._L.dsrg_seedloss_backward_dev(self.h, B, int(n_global or B), _dptr(probs), _dptr(seeds),
                                                 float(top_diff), _dptr(grad_out), _stream(stream)))
        return grad_out

    # ---- host entry points (numpy; H2D / D2H inside the call) ----
    def crf_host(self, unary, image, params, out=None, unary_layout=LAYOUT_NHWC, out_layout=LAYOUT_NHWC):
        B = image.shape[0]
        if out is None:
            shape = (B, self.H, self.W, self.M) if out_layout == LAYOUT_NHWC else (B, self.M, self.H, self.W)
            out = np.empty(shape, np.float32)
        check(self._L.dsrg_crf_batch_host(self.h, B, _hptr(unary, np.float32), unary_layout, _hptr(image, np.uint8),
                                          C.byref(params), _hptr(out, np.float32), out_layout))
        return out

    def srg_host(self, labels, probs, cues, th1, th2, renorm=False, seeds_out=None, label_map_out=None):
        B = probs.shape[0]
        if seeds_out is None:
            seeds_out = np.empty(probs.shape, np.float32)
        check(self._L.dsrg_srg_batch_host(self.h, B, _hptr(labels, np.float32), _hptr(probs, np.float32),
                                          _hptr(cues, np.float32), float(th1), float(th2), int(bool(renorm)),
                                          _hptr(seeds_out, np.float32), _hptr(label_map_out, np.int32)))
        return seeds_out

    def dsrg_forward_host(self, labels, probs, cues, image, params, th1, th2, seeds_out=None, crf_out=None):
        B = probs.shape[0]
        if seeds_out is None:
            seeds_out = np.empty(probs.shape, np.float32)
        check(self._L.dsrg_dsrg_forward_host(self.h, B, _hptr(labels, np.float32), _hptr(probs, np.float32),
                                             _hptr(cues, np.float32), _hptr(image, np.uint8), C.byref(params),
                                             float(th1), float(th2), _hptr(seeds_out, np.float32),
                                             _hptr(crf_out, np.float32)))
        return seeds_out

    # ---- SoftmaxLayer / ConstrainLossLayer (host blobs) ----
    def softmax_forward_host(self, preds):
        out = np.empty(preds.shape, np.float32)
        check(self._L.dsrg_softmax_forward_host(self.h, preds.shape[0], _hptr(preds, np.float32), _hptr(out, np.float32)))
        return out

    def softmax_backward_host(self, preds, top_diff):
        out = np.empty(preds.shape, np.float32)
        check(self._L.dsrg_softmax_backward_host(self.h, preds.shape[0], _hptr(preds, np.float32),
                                                 _hptr(top_diff, np.float32), _hptr(out, np.float32)))
        return out

    def constrainloss_forward_host(self, probs, log_smooth):
        out = np.zeros(1, np.float32)
        check(self._L.dsrg_constrainloss_forward_host(self.h, probs.shape[0], _hptr(probs, np.float32),
                                                      _hptr(log_smooth, np.float32), _hptr(out, np.float32)))
        return float(out[0])

    def constrainloss_backward_host(self, probs, log_smooth):
        gp, gl = np.empty(probs.shape, np.float32), np.empty(probs.shape, np.float32)
        check(self._L.dsrg_constrainloss_backward_host(self.h, probs.shape[0], _hptr(probs, np.float32),
                                                       _hptr(log_smooth, np.float32), _hptr(gp, np.float32),
                                                       _hptr(gl, np.float32)))
        return gp, gl

    def prepare_image_host(self, images, mean_pixel=(104.0, 117.0, 123.0), out=None):
        """(B,3,Hi,Wi) float32 network input -> (B,H,W,3) uint8 CRF image (pylayers.py:315-319 + CRF.py:32)."""
        B, _, Hi, Wi = images.shape
        if out is None:
            out = np.empty((B, self.H, self.W, 3), np.uint8)
        mean = np.asarray(mean_pixel, np.float64)
        check(self._L.dsrg_prepare_image_host(self.h, B, Hi, Wi, _hptr(images, np.float32), _hptr(mean, np.float64),
                                              _hptr(out, np.uint8)))
        return out

    def prepare_image_dev(self, images, out, mean_pixel=(104.0, 117.0, 123.0), stream=None):
        B, _, Hi, Wi = images.shape
        mean = np.asarray(mean_pixel, np.float64)
        check(self._L.dsrg_prepare_image_dev(self.h, B, Hi, Wi, _dptr(images), _hptr(mean, np.float64), _dptr(out),
                                             _stream(stream)))
        return out

    def crflayer_forward_host(self, probs, image, params, log_out=None, result=None):
        B = probs.shape[0]
        if log_out is None:
            log_out = np.empty(probs.shape, np.float32)
        check(self._L.dsrg_crflayer_forward_host(self.h, B, _hptr(probs, np.float32), _hptr(image, np.uint8),
                                                 C.byref(params), _hptr(log_out, np.float32), _hptr(result, np.float32)))
        return log_out

    def srg_last_crf_host(self, labels, cues, th1, th2, seeds_out=None):
        """SRG on the marginals of this engine's last CRF pass (one refinement, two consumers)."""
        B = cues.shape[0]
        if seeds_out is None:
            seeds_out = np.empty(cues.shape, np.float32)
        check(self._L.dsrg_srg_last_crf_host(self.h, B, _hptr(labels, np.float32), _hptr(cues, np.float32),
                                             float(th1), float(th2), _hptr(seeds_out, np.float32)))
        return seeds_out

    def crf_last_marginals_host(self, B, layout=None):
        layout = _lib.LAYOUT_NCHW if layout is None else layout
        shape = (B, self.M, self.H, self.W) if layout == _lib.LAYOUT_NCHW else (B, self.H, self.W, self.M)
        out = np.empty(shape, np.float32)
        check(self._L.dsrg_crf_last_marginals_host(self.h, B, _hptr(out, np.float32), int(layout)))
        return out

    def seedloss_forward_host(self, probs, seeds):
        """(term_bg, term_fg) local sums; loss = -(term_bg + term_fg) / N_global."""
        terms = np.zeros(2, np.float32)
        check(self._L.dsrg_seedloss_forward_host(self.h, probs.shape[0], _hptr(probs, np.float32),
                                                 _hptr(seeds, np.float32), _hptr(terms, np.float32)))
        return terms

    def seedloss_backward_host(self, probs, seeds, n_global=None, top_diff=1.0, grad=None):
        if grad is None:
            grad = np.empty(probs.shape, np.float32)
        check(self._L.dsrg_seedloss_backward_host(self.h, probs.shape[0], int(n_global or probs.shape[0]),
                                                  _hptr(probs, np.float32), _hptr(seeds, np.float32),
                                                  float(top_diff), _hptr(grad, np.float32)))
        return grad

    # ---- inference post-processing (training/tools/test-ms.py, generate_train_gt.py) ----
    def zoom_scores_host(self, scores, out=None, accumulate=False):
        """(M,h,w) float32 blob -> (H,W,M): nd.zoom(scores.transpose(1,2,0), (H/h, W/w, 1), order=1), bit-exact."""
        M, h, w = scores.shape
        if out is None:
            if accumulate:
                raise ValueError("accumulate needs `out`")
            out = np.empty((self.H, self.W, self.M), np.float32)
        check(self._L.dsrg_zoom_scores_host(self.h, _hptr(scores, np.float32), h, w, _hptr(out, np.float32),
                                            int(bool(accumulate))))
        return out

    def zoom_scores_dev(self, scores, out, accumulate=False, stream=None):
        M, h, w = scores.shape
        check(self._L.dsrg_zoom_scores_dev(self.h, _dptr(scores), h, w, _dptr(out), int(bool(accumulate)),
                                           _stream(stream)))
        return out

    @staticmethod
    def _post_args(scores, labels_sel):
        n = len(scores)
        hs = (C.c_int * n)(*[int(a.shape[1]) for a in scores])
        ws = (C.c_int * n)(*[int(a.shape[2]) for a in scores])
        sel = np.ascontiguousarray(labels_sel if labels_sel is not None else [], np.int32)
        return n, hs, ws, sel

    def predict_mask_host(self, scores, image, params=None, mode=POST_SUM_SCORES, eps=0.00001, smooth=True,
                          labels_sel=None, want_probs=False):
        """scores: list of (M,h,w) float32 blobs; image (H,W,3) uint8 -> (H,W) int32 label map
        [, (H,W,M) float32 probabilities]."""
        n, hs, ws, sel = self._post_args(scores, labels_sel)
        ptrs = (C.c_void_p * n)(*[_hptr(a, np.float32).value for a in scores])
        result = np.empty((self.H, self.W), np.int32)
        probs = np.empty((self.H, self.W, self.M), np.float32) if want_probs else None
        params = params if params is not None else crf_params()
        check(self._L.dsrg_predict_mask_host(self.h, int(mode), n, ptrs, hs, ws, _hptr(image, np.uint8), float(eps),
                                             int(bool(smooth)), C.byref(params), _hptr(sel, np.int32), int(sel.size),
                                             _hptr(result, np.int32), _hptr(probs, np.float32)))
        return (result, probs) if want_probs else result

    def predict_mask_dev(self, scores, image, result_out, params=None, mode=POST_SUM_SCORES, eps=0.00001,
                         smooth=True, labels_sel=None, probs_out=None, stream=None):
        n, hs, ws, sel = self._post_args(scores, labels_sel)
        ptrs = (C.c_void_p * n)(*[_dptr(a).value for a in scores])
        params = params if params is not None else crf_params()
        check(self._L.dsrg_predict_mask_dev(self.h, int(mode), n, ptrs, hs, ws, _dptr(image), float(eps),
                                            int(bool(smooth)), C.byref(params), _hptr(sel, np.int32), int(sel.size),
                                            _dptr(result_out), _dptr(probs_out), _stream(stream)))
        return result_out

    # ---- AnnotationLayer.forward (pylayers.py:369-387) ----
    @staticmethod
    def _annot_args(tags, cues, flip):
        """tags: per image 1-D class ids; cues: per image (3,K) int arrays (class,row,col) -> CSR int32."""
        B = len(tags)
        toff = np.zeros(B + 1, np.int32)
        coff = np.zeros(B + 1, np.int32)
        for i in range(B):
            toff[i + 1] = toff[i] + np.asarray(tags[i]).size
            coff[i + 1] = coff[i] + np.asarray(cues[i]).reshape(3, -1).shape[1]
        tg = np.ascontiguousarray(np.concatenate([np.asarray(t).reshape(-1) for t in tags]) if toff[B] else [], np.int32)
        ci = np.ascontiguousarray(np.concatenate([np.asarray(c).reshape(3, -1) for c in cues], axis=1) if coff[B]
                                  else np.zeros((3, 0)), np.int32)
        fl = None if flip is None else np.ascontiguousarray(flip, np.int32)
        return B, toff, tg, coff, ci, fl

    def annotation_forward_host(self, tags, cues, flip=None, images=None):
        """-> labels (B,1,1,M), cues (B,M,H,W)[, images (B,3,Hi,Wi)] float32."""
        B, toff, tg, coff, ci, fl = self._annot_args(tags, cues, flip)
        labels = np.empty((B, 1, 1, self.M), np.float32)
        dense = np.empty((B, self.M, self.H, self.W), np.float32)
        out_im = None if images is None else np.empty(images.shape, np.float32)
        Hi, Wi = (0, 0) if images is None else images.shape[2:]
        check(self._L.dsrg_annotation_forward_host(self.h, B, _hptr(toff, np.int32), _hptr(tg, np.int32),
                                                   _hptr(coff, np.int32), _hptr(ci, np.int32), _hptr(fl, np.int32),
                                                   _hptr(images, np.float32), Hi, Wi, _hptr(labels, np.float32),
                                                   _hptr(dense, np.float32), _hptr(out_im, np.float32)))
        return (labels, dense) if images is None else (labels, dense, out_im)

    def annotation_forward_dev(self, tags, cues, labels_out, cues_out, flip=None, images=None, images_out=None,
                               stream=None):
        B, toff, tg, coff, ci, fl = self._annot_args(tags, cues, flip)
        Hi, Wi = (0, 0) if images is None else images.shape[2:]
        check(self._L.dsrg_annotation_forward_dev(self.h, B, _hptr(toff, np.int32), _hptr(tg, np.int32),
                                                  _hptr(coff, np.int32), _hptr(ci, np.int32), _hptr(fl, np.int32),
                                                  _dptr(images), Hi, Wi, _dptr(labels_out), _dptr(cues_out),
                                                  _dptr(images_out), _stream(stream)))
        return labels_out, cues_out

    # ---- per-kernel timing ----
    def profile(self, enable):
        check(self._L.dsrg_engine_profile(self.h, int(bool(enable))))

    def profile_read(self):
        """{kernel class: (total ms, launches)} since the last read (synchronises the device)."""
        n = self._L.dsrg_profile_tag_count()
        ms = np.zeros(n, np.float32)
        cnt = np.zeros(n, np.int64)
        check(self._L.dsrg_engine_profile_read(self.h, _hptr(ms, np.float32), _hptr(cnt, np.int64)))
        return {self._L.dsrg_profile_tag_name(t).decode(): (float(ms[t]), int(cnt[t])) for t in range(n) if cnt[t]}

    # ---- introspection ----
    def lattice_sizes(self, B):
        vs = np.zeros(1, np.int32)
        vb = np.zeros(B, np.int32)
        check(self._L.dsrg_engine_lattice_sizes(self.h, B, _hptr(vs, np.int32), _hptr(vb, np.int32)))
        return int(vs[0]), vb

    def norms(self, B):
        ns = np.zeros(self.H * self.W, np.float32)
        nb = np.zeros((B, self.H * self.W), np.float32)
        check(self._L.dsrg_engine_copy_norm(self.h, 0, B, _hptr(ns, np.float32)))
        check(self._L.dsrg_engine_copy_norm(self.h, 1, B, _hptr(nb, np.float32)))
        return ns, nb


class DenseCRF(object):
    """Same surface as the reference's Cython extension type krahenbuhl2013.wrapper.DenseCRF
    (CRF/krahenbuhl2013/wrapper.pyx:20-60), backed by dsrg_densecrf_* (host pointers)."""

    def __init__(self, W, H, nlabels):
        self._L = _lib.lib()
        self.h = self._L.dsrg_densecrf_create(int(W), int(H), int(nlabels))
        if not self.h:
            raise DsrgError(_lib.E_CUDA, self._L.dsrg_last_error().decode())
        self._n = int(W) * int(H)
        self._m = int(nlabels)

    def __del__(self):
        if getattr(self, "h", None):
            self._L.dsrg_densecrf_destroy(self.h)
            self.h = None

    def set_unary_energy(self, unary_costs):
        u = np.ascontiguousarray(unary_costs, np.float32)  # float[:] memoryview in the reference
        if u.ndim != 1 or u.size != self._n * self._m:
            raise ValueError("unary_costs must be a flat float32 buffer of npixels*nlabels")
        check(self._L.dsrg_densecrf_set_unary_energy(self.h, _hptr(u, np.float32)))

    def add_pairwise_energy(self, w1, theta_alpha_1, theta_alpha_2, theta_betta_1, theta_betta_2, theta_betta_3,
                            w2, theta_gamma_1, theta_gamma_2, im):
        im = np.ascontiguousarray(im, np.uint8)  # unsigned char[:] in the reference
        if im.ndim != 1 or im.size != self._n * 3:
            raise ValueError("im must be a flat uint8 buffer of npixels*3")
        check(self._L.dsrg_densecrf_add_pairwise_energy(self.h, w1, theta_alpha_1, theta_alpha_2, theta_betta_1,
                                                        theta_betta_2, theta_betta_3, w2, theta_gamma_1,
                                                        theta_gamma_2, _hptr(im, np.uint8)))

    def map(self, n_iters=10):
        labels = np.empty(self._n, dtype=np.int32)
        check(self._L.dsrg_densecrf_map(self.h, int(n_iters), _hptr(labels, np.int32)))
        return labels

    def inference(self, n_iters=10):
        probs = np.empty(self._n * self._m, dtype=np.float32)
        check(self._L.dsrg_densecrf_inference(self.h, int(n_iters), _hptr(probs, np.float32)))
        return probs
