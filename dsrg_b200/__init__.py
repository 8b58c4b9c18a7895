"""dsrg_b200 -- B200-native (sm_100a) implementation of DSRG's per-image pixel-labelling hot path:
dense-CRF mean-field refinement -> seeded region growing -> balanced seeding loss.

Host code is Python over a C ABI (include/dsrg_b200.h, dsrg_b200/lib/libdsrg_b200.so); the CUDA
library is loaded lazily by :mod:`dsrg_b200._lib` and there is no CPU fallback.
"""
__version__ = "0.1.0"
