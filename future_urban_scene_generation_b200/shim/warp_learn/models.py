"""`from warp_learn.models import G_Resnet, get_icn_inputs` (run_test.py:21, trajectory_inference.py:26).

`G_Resnet` and `get_icn_inputs` -> the B200 implementations (`get_icn_inputs` returns its tensor on
the CUDA device, so the caller's `.to(device)` is a no-op).  Everything else the reference's module defines (the training-only
`D_NLayersMulti` / `GANLoss`) keeps coming from the reference checkout: its `warp_learn/models.py` is loaded under a private
name from the path the package `__init__` found on sys.path, and its public names are re-exported here; the reference's own
versions stay reachable as `G_Resnet_reference` / `get_icn_inputs_reference`.  Without a reference checkout only the two B200
names are available."""
import importlib.util
import os
import sys

from future_urban_scene_generation_b200.warp_learn.models import G_Resnet  # noqa: F401

_pkg = sys.modules[__package__]
_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(getattr(_pkg, "__path__", [])):
    _cand = os.path.join(_p, "models.py")
    if os.path.abspath(_p) != _here and os.path.exists(_cand):
        _spec = importlib.util.spec_from_file_location(__package__ + "._reference_models", _cand)
        _ref = importlib.util.module_from_spec(_spec)
        sys.modules[_spec.name] = _ref
        _spec.loader.exec_module(_ref)
        for _name in dir(_ref):
            if not _name.startswith("_") and _name != "G_Resnet":
                globals()[_name] = getattr(_ref, _name)
        G_Resnet_reference = _ref.G_Resnet
        get_icn_inputs_reference = _ref.get_icn_inputs
        break

from future_urban_scene_generation_b200.frame_ops import get_icn_inputs  # noqa: E402,F401  (after the re-export loop: ours wins)
