"""B200-native (sm_100a) implementation of the LoRA + T2I-Adapter conditioned SD-1.5 denoising loop.

Drop-in for the reference's adapter API (``src/adapters/{res_srdiff,modules}.py``); every FLOP runs in
``libmrisr_b200.so`` (hand-written CUDA, C ABI in ``include/mrisr_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
