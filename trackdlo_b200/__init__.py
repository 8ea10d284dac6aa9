"""trackdlo_b200: B200-native (sm_100a) implementation of TrackDLO's per-frame CPD/MCT EM
registration path (trackdlo::tracking_step / trackdlo::cpd_lle) behind a C ABI.

  csrc/          hand-written CUDA kernels + the C ABI (include/trackdlo_b200.h)
  api.py         ctypes binding used by tests and bench.py (plumbing only, no compute)
  synth.py       seeded synthetic frames
  sharding.py    frame sharding across GPUs + the single result all-gather
"""
__all__ = ["api", "synth", "sharding"]
