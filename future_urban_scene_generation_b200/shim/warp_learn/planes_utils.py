"""`from warp_learn.planes_utils import to_image, warp_unwarp_planes, get_planes, planes_to_torch`
(trajectory_inference.py:28-29, vehicle_utils.py:8, warp_learn/models.py:12) -> B200 implementation."""
from future_urban_scene_generation_b200.warp_learn.planes_utils import (  # noqa: F401
    get_planes, warp_unwarp_planes, planes_to_torch, to_image, to_image_batch)
