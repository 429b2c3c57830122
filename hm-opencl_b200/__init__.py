"""hm-opencl_b200 -- host-side Python mirror of HM-OpenCL's GPU motion-estimation interface over
libhmme_b200.so (hand-written sm_100a CUDA; C ABI in include/hmme_b200.h).

The directory name carries a hyphen (it is the reference's name), so import it through the
loader at the repository root:  `from _pkg import hm`  (or importlib, see _pkg.py).

There is no CPU path in this package: if the CUDA library is missing or no B200 is present,
construction raises.  The oracle under oracle/ is test infrastructure and is never imported here.
"""
from .api import (Group, HmmeError, HmmeLib, MotionEstimator, Plane, TEncOpenCL, NUM_CTU_PARTS, lib_path)  # noqa: F401
from .sharding import band_ctus, band_jobs, band_reference_rows, band_rows, merge_bands  # noqa: F401,E402
