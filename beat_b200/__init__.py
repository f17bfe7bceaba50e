"""
beat_b200 -- B200-native (sm_100a) implementation of BEAT's per-chain forward model + log-likelihood hot path
(fast-sweeping rupture times -> GF-library stacking -> residual -> covariance-weighted misfit), batched over the
chains of the SMC / PT samplers.  See DESIGN.md for the scope contract and INTEGRATION.md for the reference-side
binding.  Importing the package does not load CUDA; the first GPU call loads ``libbeatgpu.so`` and fails loudly
if it is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"

from . import covariance, synthetic  # noqa: F401  (pure numpy, set-up time only)


def build_library(verbose=False):
    """Compile ``libbeatgpu.so`` in-tree with nvcc for sm_100a."""
    from .build import build
    return build(verbose=verbose)
