"""`from vunet.models import Vunet_fix_res` (run_test.py:20) -> B200 implementation."""
from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res  # noqa: F401
