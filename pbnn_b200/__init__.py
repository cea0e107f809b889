"""pbnn_b200 -- B200-native (sm_100a) drop-in for pbnn's data-parallel hot path: many independent SG-MCMC
chains / deep-ensemble members over small MLPs.  See DESIGN.md; the C ABI is include/pbnn_b200.h."""
from . import logprobs, map_estimation, mcmc, models, pytree, utils  # noqa: F401
from .deep_ensembles import deep_ensembles_fn  # noqa: F401
from .models import MLP  # noqa: F401

__version__ = "0.1.0"
