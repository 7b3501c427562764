"""bonnie-32_b200 — B200-native (sm_100a CUDA) rasterizer hot path of EBonura/bonnie-32.

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of include/b32_raster.h), `abi.py`
(ctypes binding), `raster.py` (host mirror of the reference's render_mesh_15 / Framebuffer
interface), `scenes.py` (the BASELINE.json workloads).  The directory name is not a Python
identifier; import it as `bonnie32_b200` through `__graft_entry__.load_package()`.
"""
from . import abi, raster, scenes  # noqa: F401
from .abi import B32Error  # noqa: F401
from .raster import (Camera, Context, Framebuffer, Light, Mesh, RasterSettings, Texture, Texture15,  # noqa: F401
                     render_mesh, render_mesh_15)
