"""Drop-in for the reference's Cython extension ``krahenbuhl2013.wrapper``
(CRF/krahenbuhl2013/wrapper.pyx:20-60): the ``DenseCRF(W, H, nlabels)`` type with
``set_unary_energy / add_pairwise_energy / map / inference``, backed by the B200 C ABI
(dsrg_densecrf_*, include/dsrg_b200.h) instead of the CPU DenseCRFWrapper."""
from dsrg_b200.api import DenseCRF  # noqa: F401

__all__ = ["DenseCRF"]
