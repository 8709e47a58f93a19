"""xtp_b200: B200-native (sm_100a CUDA, FP64) GW-BSE tensor-contraction path of
VOTCA-XTP behind the reference's class names.  The compute lives in
``libxtpb200.so`` (C ABI in ``include/xtpb200/xtpb200.h``); this package is the
thin host-side mirror of the reference interface.  There is no CPU fallback:
every compute entry point raises if the CUDA library cannot be loaded."""

__all__ = ["synth"]
