"""emap_b200 -- B200-native (sm_100a) implementation of EMAP's volume-rendering hot path.

Drop-in surface (same names / kwargs / return dicts as the reference's src/models):
    emap_b200.udf_model.UDFNetwork, SingleVarianceNetwork, BetaNetwork
    emap_b200.udf_renderer_blending.UDFRendererBlending
Everything executes in hand-written CUDA kernels behind the C ABI of include/emap_b200.h.
"""
__version__ = "0.1.0"
