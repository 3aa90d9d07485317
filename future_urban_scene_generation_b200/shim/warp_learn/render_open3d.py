"""`from warp_learn.render_open3d import get_rendered` (warp_learn/vehicle_utils.py:9) -> the B200 rasteriser: no Open3D
window, no OpenGL context; same signature and return types (render_open3d.py:29-50)."""
from future_urban_scene_generation_b200.warp_learn.render import get_rendered, render_normals_batch  # noqa: F401
