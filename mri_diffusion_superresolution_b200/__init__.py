"""B200-native (sm_100a) implementation of the LoRA + T2I-Adapter conditioned SD-1.5 denoising loop.

Drop-in for the reference's adapter API (``src/adapters/{res_srdiff,modules}.py``); every FLOP runs in
``libmrisr_b200.so`` (hand-written CUDA, C ABI in ``include/mrisr_b200.h``).  There is no CPU fallback.

Modules: ``res_srdiff`` (the reference's function API), ``unet`` / ``adapter`` / ``controlnet`` / ``vae`` (network drop-ins),
``scheduler`` / ``sampler`` (bookkeeping, batched CUDA-graph loop), ``pipeline`` (volume in, scored slices out),
``evalmetrics`` / ``slices`` / ``mnist`` (evaluation, data preparation, the MNIST toy's runnable cells), ``ops`` / ``_lib``
(tensor-level wrappers over the C ABI), ``parallel`` (slice sharding), ``synthetic`` (random-init weights, phantom volumes).
"""
__version__ = "0.1.0"
