"""Drop-in replacement of the reference's `models.*` modules.

The reference resolves `models.model_tcnn`, `models.nerf_tcnn`, `models.ray_sampling`,
`models.rendering_tcnn` and `models.losses` through `sys.path.append(PROJECT_ROOT + "/src")`
(/root/reference/examples/run_loner.py:39-40).  `install()` puts this directory in front, so that
`src/mapping/optimizer.py`, `src/loner.py` and `analysis/*` import the B200 kernels unchanged.
"""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        del sys.modules[name]
    return here
