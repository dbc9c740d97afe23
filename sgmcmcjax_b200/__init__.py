"""sgmcmcjax_b200 - B200-native (sm_100a) engine for the SGMCMCJax hot path.

Mirrors the reference's API for that path only:
``samplers.build_*_sampler``, ``kernels.build_*_kernel``, ``ksd.imq_KSD``.
"""
__version__ = "0.1.0"
