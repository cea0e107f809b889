"""``pbnn.mcmc`` mirror (src/pbnn/mcmc/__init__.py:10-22) -- the samplers on the CUDA hot path."""
from . import kernels
from ._common import predictive_moments
from .hamiltonian import hmc, sghmc
from .langevin import pSGLD, scheduled_sgld, sgld

__all__ = ["sgld", "pSGLD", "scheduled_sgld", "sghmc", "hmc", "kernels", "predictive_moments"]
