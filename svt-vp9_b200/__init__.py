"""svt-vp9_b200: B200 (sm_100a) implementation of SVT-VP9's per-block encode hot path behind a C ABI.

  include/svtb200.h        the C ABI (drop-in boundary)
  csrc/                    CUDA kernels + ABI implementation  -> libsvtvp9_b200.so
  shim/                    reference-side binding (compiled against the reference's headers only)
  api.py                   ctypes mirror of the ABI for tests / bench
  frames.py, presets.py    picture preparation and resolved parameter sets for the BASELINE configs
  synth.py                 deterministic synthetic clips
"""
