"""`warp_learn` as the reference's callers import it.  `online_visibility` and `planes_utils` are
served by the B200 path; every other submodule (`models` with G_Resnet / get_icn_inputs,
`vehicle_utils`, `render_open3d`) keeps resolving to the reference checkout, which is found on
sys.path and appended to this package's __path__."""
import os
import sys

for _p in sys.path:
    _cand = os.path.join(_p, "warp_learn")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)) \
            and os.path.exists(os.path.join(_cand, "vehicle_utils.py")):
        __path__.append(_cand)
        break
