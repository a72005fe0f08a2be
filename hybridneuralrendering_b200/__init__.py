"""hybridneuralrendering_b200 -- B200-native per-ray sample pipeline of HybridNeuralRendering.

Host-side mirror of the reference's module API over libhnr.so (hand-written sm_100a CUDA behind a
C ABI, include/hnr.h).  No CPU / PyTorch fallback: the kernels are the only implementation.
"""
from .options import make_opt  # noqa: F401

__all__ = ["make_opt", "lighting_fast_querier", "NeuralPoints", "PointAggregator", "NeuralPointsRayMarching", "ray_march",
           "blur_update_output", "learnable_blur_update_output", "fill_invalid"]


def __getattr__(name):
    # lazy: importing the package must not require torch.cuda or the built library
    if name == "lighting_fast_querier":
        from .querier import lighting_fast_querier as v
    elif name == "NeuralPoints":
        from .neural_points import NeuralPoints as v
    elif name == "PointAggregator":
        from .point_aggregators import PointAggregator as v
    elif name in ("NeuralPointsRayMarching", "fill_invalid"):
        from . import neural_points_volumetric_model as m
        v = getattr(m, name)
    elif name == "ray_march":
        from .diff_ray_marching import ray_march as v
    elif name == "blur_update_output":
        from .blur import blur_update_output as v
    elif name == "learnable_blur_update_output":
        from .blur import learnable_blur_update_output as v
    else:
        raise AttributeError(name)
    return v
