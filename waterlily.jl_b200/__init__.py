"""waterlily.jl_b200 — host-side mirror of WaterLily.jl's `Simulation`/`sim_step!`/`Flow`/`MultiLevelPoisson`
interface over the B200 C-ABI library (csrc/libwl_b200.so, declared in include/wl_b200.h).

The directory name contains a dot, so import it through the root-level shim:  `import wl_b200`.
Everything that computes runs in hand-written sm_100a kernels; there is no CPU fallback and this
package never touches oracle/.
"""
from .lib import load_library, build_library, WLError, library_path, release_pool  # noqa: F401
from .body import AutoBody, NoBody, Sphere, Torus, measure_body, mu0_kernel, mu1_kernel  # noqa: F401
from .flow import Flow, MultiLevelPoisson, Poisson, mom_step, quick, cds, vanLeer, loc_grid, dist_unique_id, Forcing, TimeBC, sgs, smagorinsky  # noqa: F401
from .simulation import Simulation, sim_step, sim_time, measure, sim_info  # noqa: F401
from .metrics import (pressure_force, viscous_force, total_force, pressure_moment, viscous_moment, total_moment, MeanFlow,  # noqa: F401
                      save, load)
from . import metrics  # noqa: F401
