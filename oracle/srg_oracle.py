"""oracle/srg_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracles for the seeded-region-growing step ``generate_seed_step``
(pylayers/pylayers/pylayers.py:237-275) and its connected-component labeller
(pylayers/pylayers/CC_labeling_8.py:103-282).

Three flavours, all returning the same (C,H,W) float32 0/1 seed array:

* :func:`reference_generate_seed_step` -- the reference's OWN functions, executed in place
  from /root/reference (dev container only; nothing is copied: the FunctionDef is pulled out
  of the module with ``ast`` because the module itself imports caffe/theano/cPickle).  This
  is what pins the other two (tests/golden/make_golden.py freezes its outputs as fixtures).
* :func:`srg_closed_form` -- the vectorised closed form of SURVEY.md 3.3 (numpy + an
  equal-label 8-connectivity labelling): fast enough for 321x321 / 513x513 sweeps.
* :func:`srg_faithful` -- a pure-Python restatement that follows the reference's loop
  structure step by step, INCLUDING the two-pass labeller's unused region-adjacency
  bookkeeping (CC_labeling_8.py:180-190, :201-207), so that its run time is representative
  of the reference's CPU cost.  Used for small parity cases and as bench.py's
  ``cpu_baseline`` (kind "port") on the GPU box, where /root/reference does not exist.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this module.
"""
import ast
import os
import sys

import numpy as np

REF_ROOT = "/root/reference"
_REF_FN = None


def reference_available():
    return os.path.exists(os.path.join(REF_ROOT, "pylayers/pylayers/pylayers.py"))


def reference_generate_seed_step():
    """Return the reference's own ``generate_seed_step`` bound to its own CC_labeling_8."""
    global _REF_FN
    if _REF_FN is None:
        pdir = os.path.join(REF_ROOT, "pylayers/pylayers")
        if pdir not in sys.path:
            sys.path.insert(0, pdir)
        import CC_labeling_8  # the reference's module, imported in place
        with open(os.path.join(pdir, "pylayers.py")) as fh:
            tree = ast.parse(fh.read())
        fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "generate_seed_step"]
        assert len(fn) == 1
        mod = ast.Module(body=fn, type_ignores=[])
        ns = {"np": np, "CC_labeling_8": CC_labeling_8}
        exec(compile(mod, os.path.join(pdir, "pylayers.py"), "exec"), ns)
        _REF_FN = ns["generate_seed_step"]
    return _REF_FN


def run_reference(labels, cues, probs, th1, th2):
    """labels (C,), cues (C,H,W) f32 0/1, probs (C,H,W) -> upcast to float64 like the
    reference's data (pylayers.py:323-330 builds a float64 array)."""
    fn = reference_generate_seed_step()
    seed_c = np.array(cues, np.float32, copy=True)
    return fn([np.asarray(labels), seed_c, np.asarray(probs, np.float64), th1, th2])


# --------------------------------------------------------------------------------------
# label map: pylayers.py:240-257
# --------------------------------------------------------------------------------------
def label_map_closed_form(labels, cues, probs, th1, th2):
    """(H,W) int32: 0 = unlabelled, c+1 = class c."""
    labels = np.asarray(labels)
    probs = np.asarray(probs, np.float64)
    cls_index = np.where(labels == 1)[0]                      # :240
    sel = probs[cls_index]                                    # :241
    probs_c = np.argmax(sel, axis=0)                          # :242 first max wins
    probs_p = np.max(sel, axis=0)                             # :243
    C, H, W = cues.shape
    lm = np.zeros((H, W), np.int32)
    for c in range(C):                                        # :248-250 highest class index wins
        lm[cues[c] > 0] = c + 1
    cmap = cls_index[probs_c]
    take = (probs_p > th2) & ((cmap != 0) | (probs_p > th1))   # :253-257
    lm[take] = cmap[take] + 1
    return lm


def _label_components_8(lm):
    """Equal-label 8-connectivity components of the non-zero pixels of ``lm``.
    Returns (H,W) int64 root ids (pixel index of the component's minimum raster index),
    -1 where lm == 0.  Plain union-find written for clarity, vectorised edge lists."""
    H, W = lm.shape
    idx = np.arange(H * W, dtype=np.int64).reshape(H, W)
    parent = np.arange(H * W, dtype=np.int64)
    fg = lm > 0
    edges = []
    for dy, dx in ((0, 1), (1, 0), (1, 1), (1, -1)):
        a = (slice(0, H - dy), slice(max(0, -dx), W - max(0, dx)))
        b = (slice(dy, H), slice(max(0, dx), W - max(0, -dx)))
        same = fg[a] & (lm[a] == lm[b])
        edges.append(np.stack([idx[a][same], idx[b][same]], 1))
    edges = np.concatenate(edges, 0)
    # iterate hooking + pointer jumping (min-label propagation) until stable
    while True:
        pa, pb = parent[edges[:, 0]], parent[edges[:, 1]]
        lo, hi = np.minimum(pa, pb), np.maximum(pa, pb)
        changed = lo != hi
        if not changed.any():
            break
        np.minimum.at(parent, hi[changed], lo[changed])
        while True:
            gp = parent[parent]
            if (gp == parent).all():
                break
            parent = gp
    root = parent.reshape(H, W).copy()
    root[~fg] = -1
    return root


def srg_closed_form(labels, cues, probs, th1, th2, return_label_map=False):
    """Vectorised closed form of generate_seed_step."""
    labels = np.asarray(labels)
    cues = np.asarray(cues, np.float32)
    C, H, W = cues.shape
    lm = label_map_closed_form(labels, cues, probs, th1, th2)
    root = _label_components_8(lm)
    cls = np.clip(lm - 1, 0, C - 1)
    own = np.take_along_axis(cues, cls[None], 0)[0] == 1       # seed_c[c,x,y] == 1  (:266)
    own &= lm > 0
    present = np.zeros(C + 1, bool)
    present[1:] = labels == 1
    own &= present[lm]                                         # only present classes grow (:259)
    nseed = np.sum(cues, axis=0)                               # np.sum(seed_c[:,x,y]) (:268)
    excl = (~own) & (nseed == 1)
    hc = np.zeros(H * W, bool)
    hc[root[own]] = True                                       # high_confidence_set_label
    grow = (lm > 0) & present[lm] & hc[np.where(root >= 0, root, 0)] & (root >= 0) & ~excl
    out = cues.copy()
    ys, xs = np.nonzero(grow)
    out[lm[ys, xs] - 1, ys, xs] = 1                            # :271-273
    if return_label_map:
        return out, lm
    return out


# --------------------------------------------------------------------------------------
# faithful pure-Python restatement (time-representative)
# --------------------------------------------------------------------------------------
class _Forest(object):
    """Union by rank + path compression; CC_labeling_8.py:51-84."""

    def __init__(self, n):
        self.parent = [0] * n
        self.rank = [0] * n
        self.neighbors = [[]] * n          # the unused adjacency lists (:58)

    def make(self, x):
        self.parent[x] = x
        self.rank[x] = 0

    def find(self, x):
        if self.parent[x] != x:
            self.parent[x] = self.find(self.parent[x])
        return self.parent[x]

    def union(self, x, y):
        xr, yr = self.find(x), self.find(y)
        if xr == yr:
            return
        if self.rank[xr] < self.rank[yr]:
            self.parent[xr] = yr
        elif self.rank[xr] > self.rank[yr]:
            self.parent[yr] = xr
        else:
            self.parent[yr] = xr
            self.rank[xr] += 1


def _two_pass_label(mat):
    """Two-pass 8-connectivity labelling of EQUAL-VALUED regions of a 0/1 matrix
    (both the 1- and the 0-regions get labels), CC_labeling_8.py:112-197, with the
    W / N / NW / NE neighbour test of :244-282 and the adjacency side effect of :180-190."""
    R, Cn = len(mat), len(mat[0])
    lab = [[0] * Cn for _ in range(R)]
    forest = _Forest(R * Cn)
    nxt = 0

    def touch(i, j, k, l):                 # CC_labeling_8.py:201-207 (result unused by the caller)
        a, b = lab[i][j], lab[k][l]
        forest.neighbors[a] = np.unique([x for x in forest.neighbors[a]] + [b])
        forest.neighbors[b] = np.unique([x for x in forest.neighbors[b]] + [a])

    for i in range(R):
        for j in range(Cn):
            v = mat[i][j]
            w = j > 0 and mat[i][j - 1] == v
            n = i > 0 and mat[i - 1][j] == v
            nw = i > 0 and j > 0 and mat[i - 1][j - 1] == v
            ne = i > 0 and j < Cn - 1 and mat[i - 1][j + 1] == v
            got = []
            if w:
                got.append(lab[i][j - 1])
            if n:
                got.append(lab[i - 1][j])
            if nw:
                got.append(lab[i - 1][j - 1])
            if ne:
                got.append(lab[i - 1][j + 1])
            if not got:
                lab[i][j] = nxt
                forest.make(nxt)
                nxt += 1
            elif len(got) > 1:
                lab[i][j] = min(got)
                for t in range(len(got) - 1):
                    forest.union(got[t], got[t + 1])
            else:
                lab[i][j] = got[0]
            # the reference's index names are swapped relative to the direction names
            # (neighbors[0] is the (i, j-1) pixel but the update touches (i-1, j), etc.)
            if not w and i > 0:
                touch(i, j, i - 1, j)
            if not n and j > 0:
                touch(i, j, i, j - 1)
            if not nw and i > 0 and j > 0:
                touch(i, j, i - 1, j - 1)
            if not ne and i > 0 and j < Cn - 1:
                touch(i, j, i - 1, j + 1)
    for i in range(R):
        for j in range(Cn):
            lab[i][j] = forest.find(lab[i][j])
    return lab


def srg_faithful(labels, cues, probs, th1, th2):
    """Loop-for-loop restatement of generate_seed_step, pylayers.py:237-275."""
    labels = np.asarray(labels)
    seed_c = np.array(cues, np.float32, copy=True)
    probs = np.asarray(probs, np.float64)
    cls_index = np.where(labels == 1)[0]
    sel = probs[cls_index]
    probs_c = np.argmax(sel, axis=0)
    probs_p = np.max(sel, axis=0)
    C, H, W = seed_c.shape
    label_map = np.zeros((H, W))
    i0, i1, i2 = np.where(seed_c > 0)
    label_map[i1, i2] = i0 + 1
    for (x, y), value in np.ndenumerate(probs_p):
        c = cls_index[probs_c[x, y]]
        if value > th2:
            if not c == 0:
                label_map[x, y] = c + 1
            elif value > th1:
                label_map[x, y] = c + 1
    for c in cls_index:
        mat = (label_map == (c + 1)).astype(int)
        lab = _two_pass_label(mat)
        keep = set()
        for (x, y), value in np.ndenumerate(mat):
            if value == 1 and seed_c[c, x, y] == 1:
                keep.add(lab[x][y])
            elif value == 1 and np.sum(seed_c[:, x, y]) == 1:
                lab[x][y] = -1
        for (x, y), value in np.ndenumerate(np.array(lab)):
            if value in keep:
                seed_c[c, x, y] = 1
    return seed_c
