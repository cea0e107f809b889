"""``pbnn.mcmc`` mirror (src/pbnn/mcmc/__init__.py:10-22) -- the samplers on the CUDA hot path."""
from . import kernels
from ._common import predictive_moments
from .hamiltonian import hmc, sghmc, sghmc_cv, sghmc_svrg
from .langevin import cyclical_sgld, pSGLD, scheduled_sgld, sgld, sgld_cv, sgld_svrg

__all__ = ["sgld", "pSGLD", "scheduled_sgld", "sgld_cv", "sgld_svrg", "cyclical_sgld", "sghmc", "sghmc_cv",
           "sghmc_svrg", "hmc", "kernels", "predictive_moments"]
