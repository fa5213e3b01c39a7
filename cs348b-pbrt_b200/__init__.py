"""cs348b-pbrt_b200: B200-native volumetric photon-mapping path of piwell/CS348B-pbrt.

The product is csrc/libpv.so (hand-written sm_100a CUDA behind the C ABI of
include/pv.h).  This package is the thin Python host mirror used by tests and
bench.py: it loads the library with ctypes and offers `PhotonVolume`, whose
methods carry the reference's names (Preprocess / Li / Transmittance,
core/photonshooter.h:81-116, integrators/photonvolume.h:14-38).
There is no CPU fallback: if libpv.so is missing or no CUDA device is present
the calls raise.
"""
from . import _abi, sceneio  # noqa: F401
from .api import PhotonVolume, PVError, load_library, library_path  # noqa: F401
