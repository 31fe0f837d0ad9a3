"""pzflow_b200 -- B200-native evaluator for pzflow's neural-spline-flow hot path.

Drop-in for ``Flow.log_prob`` / ``Flow.sample`` / ``Flow.posterior`` (and the
``FlowEnsemble`` forms) behind the reference's own API: same ``Flow`` /
``FlowEnsemble`` signatures, ``Bijector_Info`` tuples and params pytrees; the
arithmetic runs in hand-written sm_100a kernels reached through a C-ABI
(``include/pzflow_b200.h``, ``pzflow_b200/libpzflow_b200.so``).
"""
from .flow import Flow
from .flowEnsemble import FlowEnsemble

__version__ = "0.1.0"
__all__ = ["Flow", "FlowEnsemble"]
