from .cameras import PerspectiveCameras, CamerasBase, look_at_view_transform  # noqa: F401
from typing import NamedTuple
import torch


class RayBundle(NamedTuple):
    origins: torch.Tensor
    directions: torch.Tensor
    lengths: torch.Tensor
    xys: torch.Tensor


class GridRaysampler:  # import-only stub
    def __init__(self, *a, **k):
        pass


def ray_bundle_to_ray_points(ray_bundle):
    """points = origins[..., None, :] + lengths[..., :, None] * directions[..., None, :]"""
    return ray_bundle.origins[..., None, :] + ray_bundle.lengths[..., :, None] * ray_bundle.directions[..., None, :]
