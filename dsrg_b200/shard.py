"""Multi-GPU host logic: one process per GPU, images shard by index, no data-path collective.

Every image of the DSRG pass is independent (pylayers/pylayers/pylayers.py:325-326, :341-342), so a
batch is split contiguously over the ranks and each rank runs its own engine.  The
only exchange on the path is the balanced seeding loss, a mean over the GLOBAL batch
(pylayers.py:136-139): two partial sums (+ the local image count) are all-reduced with
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_images, rank, world):
    """Contiguous [lo, hi) of image indices owned by ``rank``: image i -> rank floor(i*world/n)."""
    if not (0 <= rank < world) or n_images < 0:
        raise ValueError("bad shard request")
    lo = -(-rank * n_images // world)          # ceil(rank*n/world)
    hi = -(-(rank + 1) * n_images // world)
    return lo, hi


def world():
    """(rank, world_size) of the initialised default process group, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def allreduce_loss_terms(terms, n_local):
    """SUM-all-reduce (term_bg, term_fg, n_local).  Returns (global terms or None, global N).
    ``terms`` may be None when only the global batch size is needed (backward)."""
    rank, ws = world()
    if ws == 1:
        return terms, int(n_local)
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = (0.0, 0.0) if terms is None else (float(terms[0]), float(terms[1]))
    buf = torch.tensor([t[0], t[1], float(n_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(buf)
    out = buf.cpu().numpy()
    return (None if terms is None else np.array(out[:2])), int(round(out[2]))


def balanced_loss_from_terms(terms, n_global):
    """loss = -(sum_n S_bg/cnt_bg + sum_n S_fg/cnt_fg) / N  (pylayers.py:136-139)."""
    return -(float(terms[0]) + float(terms[1])) / float(n_global)
