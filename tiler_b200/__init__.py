"""tiler_b200 -- B200-native data-parallel core of the TileMotion encoder (gligli/tiler), behind the C ABI of
libtm_gpu.so.  See DESIGN.md for the path and its boundary, INTEGRATION.md for the FreePascal bindings."""
from . import api, synth  # noqa: F401
from .api import *  # noqa: F401,F403
