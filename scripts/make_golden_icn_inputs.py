"""Goldens for the ICN input packing (warp_learn/models.py:323-366 get_icn_inputs), produced with the imported reference
function itself (cv2.resize, cv2.cvtColor, PIL, torchvision): sha1 of gen_in + crop_info per synthetic vehicle, and a check
that oracle/frame_oracle.py reproduces them bit for bit.  Appends the "icn_inputs" section to tests/golden/frame_golden.json.
Only runs where /root/reference, cv2, PIL and torchvision exist (the build container)."""
import hashlib
import json
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")
import warnings
warnings.filterwarnings("ignore")
import numpy as np

from warp_learn.models import get_icn_inputs                      # the reference
from future_urban_scene_generation_b200 import synth
from oracle import frame_oracle as FO

cases = []
for idx, hw in [(0, (360, 640)), (1, (360, 640)), (3, (360, 640)), (4, (256, 256)), (7, (720, 1280)), (8, (360, 640)), (11, (360, 640)), (12, (300, 300))]:
    planes, normal, mask, central = synth.make_icn_pack_case(idx, hw)
    ref, info = get_icn_inputs(planes.copy(), normal.copy(), mask.copy(), central.copy(), 256, 256)
    ref = ref.numpy()
    got, ginfo = FO.get_icn_inputs(planes, normal, mask, central)
    assert ref.shape == (1, 21, 256, 256) and np.array_equal(ref[0].view(np.int32), got.view(np.int32)), ("oracle != reference", idx)
    info = {k: [int(v) for v in info[k]] for k in ("crop_xy_min", "pad_xy_before", "pad_xy_after", "crop_size_orig")}
    assert all(list(ginfo[k]) == info[k] for k in info), (idx, ginfo, info)
    cases.append({"idx": idx, "frame_hw": list(hw), "sha1": hashlib.sha1(np.ascontiguousarray(ref[0]).tobytes()).hexdigest(), **info,
                  "mean": float(ref.astype(np.float64).mean())})
    print(idx, hw, "oracle == reference get_icn_inputs (bit-exact)", info["crop_size_orig"])
path = os.path.join(ROOT, "tests", "golden", "frame_golden.json")
gold = json.load(open(path))
gold["icn_inputs"] = cases
json.dump(gold, open(path, "w"), indent=1)
print("updated", path)
