"""`from warp_learn.online_visibility import pascal_texture_planes, compute_visibility`
(trajectory_inference.py:27, warp_learn/vehicle_utils.py:7) -> B200 implementation."""
from future_urban_scene_generation_b200.warp_learn.online_visibility import (  # noqa: F401
    pascal_texture_planes, compute_visibility, compute_visibility_batch)
