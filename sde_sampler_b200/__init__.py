"""B200-native fused rollout for juliusberner/sde_sampler's optimal-control losses.

Public surface = the reference's loss plug-in interface (sde_sampler/losses/oc.py):
`FusedTimeReversalLoss`, `FusedReferenceSDELoss`, `FusedExponentialIntegratorSDELoss`, `Results`.
Importing the losses requires the in-tree CUDA library (`python -m sde_sampler_b200.build`).
"""
from .integrator import FusedEulerIntegrator  # noqa: F401
from .trainer import FusedAdamEMA, eval_moments, sample_gauss_prior  # noqa: F401
from .losses import (FusedExponentialIntegratorSDELoss, FusedOCLoss, FusedReferenceSDELoss,  # noqa: F401
                     FusedTimeReversalLoss, Results)

__all__ = ["FusedTimeReversalLoss", "FusedReferenceSDELoss", "FusedExponentialIntegratorSDELoss",
           "FusedOCLoss", "Results", "FusedEulerIntegrator", "FusedAdamEMA", "sample_gauss_prior", "eval_moments"]
