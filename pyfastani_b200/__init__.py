"""pyfastani_b200 -- the pyfastani mapping path (Sketch -> Mapper -> Hit) on NVIDIA B200.

Same Python surface as ``pyfastani`` (src/pyfastani/__init__.py:2-11 in the reference); all
computation runs in ``lib/libfastani_b200.so`` (CUDA, sm_100a).  There is no CPU fallback: if
the compiled extension is missing, importing this package fails.
"""
from . import _fastani
from ._fastani import (
    MAX_KMER_SIZE,
    Communicator,
    CudaError,
    DeviceFasta,
    DeviceSequence,
    Hit,
    Mapper,
    MinimizerIndex,
    MinimizerInfo,
    Minimizers,
    PackedSequence,
    Position,
    Sketch,
    device_count,
)

__version__ = _fastani.__version__
__all__ = [
    "MAX_KMER_SIZE", "Hit", "Mapper", "MinimizerIndex", "MinimizerInfo", "Minimizers", "Position", "Sketch",
    "Communicator", "CudaError", "DeviceFasta", "DeviceSequence", "PackedSequence", "device_count",
]
