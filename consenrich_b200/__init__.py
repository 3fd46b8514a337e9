"""consenrich_b200 -- B200 (sm_100a) implementation of Consenrich's state-space hot path.

Only what the path needs lives here:

* ``csrc/``      hand-written CUDA kernels + the C ABI (``include/consenrich_b200.h``)
* ``_lib``       ctypes binding of ``lib/libconsenrich_b200.so``
* ``native``     drop-in replacements for the six hot-path functions of ``consenrich.cconsenrich`` (and the
                 three background-track functions that call them from the other side)
* ``driver``     device versions of the driver-side ``[tracks x intervals]`` reductions of ``consenrich.core``
* ``device``     device-resident sweeps on torch tensors (torch is only a tensor carrier)
* ``sharding``   chromosome / bin-range sharding across ranks (torch.distributed plumbing)
* ``writers``    bedGraph chunks formatted on the device, byte-identical to the reference's pandas writer
* ``bigwig``     host-side bigWig container for the finished bedGraph files (what the reference asks pyBigWig for)

There is no CPU implementation: importing works anywhere, calling requires the built library
and a CUDA device, and fails loudly otherwise.
"""
from . import _lib  # noqa: F401
from .native import (cEMA, cFinalizeMuncEBTrack, cMuncObservationMomentSeedPass,  # noqa: F401
                     cMuncSmoothDenseLocalEvidence, cbackgroundWeightedStats,
                     cbackgroundWeightedStatsWithSupport,
                     cbackwardPass, cbackwardPassLevel, cfixedBackgroundECM, cfixedBackgroundECMLevel,
                     cforwardPass, cforwardPassLevel, csolveZeroCenteredBackground, install, sweep, uninstall)

from .driver import install_driver, uninstall_driver  # noqa: E402,F401
from . import writers  # noqa: E402,F401
from . import bigwig  # noqa: E402,F401

__version__ = "0.1.0"
