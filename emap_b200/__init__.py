"""emap_b200 -- B200-native (sm_100a) implementation of EMAP's volume-rendering hot path.

Drop-in surface (same names / kwargs / return dicts as the reference's src/models):
    emap_b200.udf_model.UDFNetwork, SingleVarianceNetwork, BetaNetwork
    emap_b200.udf_renderer_blending.UDFRendererBlending
    emap_b200.udf_model.RenderingNetwork                     (standalone operator; dead code in the reference)
    emap_b200.extract_pointcloud.get_pointcloud_from_udf, emap_b200.ray_sampler.RaySampler
Additions that have no counterpart in the (single-GPU, eager) reference:
    emap_b200.parallel   rays shard across ranks, one in-place flat-buffer gradient all-reduce
    emap_b200.graph      GraphedStep: a whole iteration captured in one CUDA graph
Everything executes in hand-written CUDA kernels behind the C ABI of include/emap_b200.h.
"""
__version__ = "0.2.0"
