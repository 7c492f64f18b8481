"""xnrs_b200 — B200-native (sm_100a) implementation of the xnrs bi-encoder hot path.

Drop-in for the reference's ``xnrs.models`` modules / ``make_model`` factory and the trainer loss hooks;
all tensor math runs in the hand-written CUDA kernels of ``xnrs_b200/csrc`` behind the C ABI declared in
``include/xnrs_b200.h``.  There is no CPU fallback: importing works anywhere, but every op raises unless
the compiled library is present and its inputs live on a CUDA device.
"""
__version__ = '0.1.0'
