"""dabry_b200 - B200-native extremal-field solver behind the Python API of dabry (dabry-route/dabry v1.1).

Drop-in surface: `NavigationProblem`, `FlowField` / `DiscreteFF`, `Dynamics`, `SolverEFSimple`,
`SolverEFResampling`, `SolverEFTrimming`.  The solvers run on hand-written CUDA kernels (sm_100a) through the
C ABI of include/dabry_b200.h; there is no CPU execution path.
"""
__version__ = '0.1.0'

from dabry_b200.misc import Coords, Units, Utils, Chrono  # noqa: F401
from dabry_b200.flowfield import FlowField, DiscreteFF, WrapperFF  # noqa: F401
from dabry_b200.problem import NavigationProblem  # noqa: F401
from dabry_b200.solver_ef import (SolverEF, SolverEFSimple, SolverEFResampling, SolverEFTrimming,  # noqa: F401
                                  Site, SiteManager, ClosureReason, NeuteringReason)
