"""ofq_b200 — B200-native (sm_100a) implementation of the OFQ quantization-aware-training hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C-ABI (include/ofq_b200.h) -> libofq_b200.so
  ops.py           ctypes/tensor wrappers over the C-ABI
  quantization/    host-side mirror of the reference's `src/quantization` package (same class names,
                   constructor arguments, forward signatures and state-dict keys)
  host/            timm-free DeiT / Swin host models (LayerNorm, residuals, heads) used by tests and bench
  cga.py           CGA-masked AdamW optimizer (cga.py:953-1013 semantics in one fused kernel)
"""
__version__ = "0.1.0"
