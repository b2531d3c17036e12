"""videovanish_b200 - B200-native pixel pipeline behind VideoVanish's
``diffuerase.run_infill_on_frames`` / ``tools.*`` call surface.

The compute path is hand-written sm_100a CUDA in ``csrc/`` behind a C-ABI shared
library (``include/vvb200.h``); this package is the Python host side that mirrors
the reference's interface.  There is NO CPU fallback: importing ``_lib`` raises if
the library is missing, and every op raises if it is called without a CUDA device.
"""
__version__ = "0.1.0"
