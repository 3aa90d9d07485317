"""Golden vectors for the paste-back row, produced with cv2 and the imported reference (/root/reference):
tests/golden/frame_golden.json

  * resize: sha1 of cv2.resize(src, dsize) for seeded random uint8 images over a list of (src, dst) size pairs
    (incl. identity, exact halving, 256 -> vehicle-sized, strong up/down-scaling, tiny images)
  * crop_info: utils/crop_utils.py square_crop_from_bbox geometry for synthetic boxes (synth.make_paste_case)
  * pack: sha1 of x / y_tilde from the reference lines trajectory_inference.py:205-227 (cv2.resize, the reference's
    square_crop_from_bbox and to_tensor, torch cat / interpolate) on synthetic vehicles (synth.make_pack_case)
  * paste: sha1 of the frame after the five reference lines (trajectory_inference.py:236-250) executed with cv2.resize
    for a sequence of vehicles pasted into one frame in order

Only runs where /root/reference and cv2 exist (the build container)."""
import hashlib
import json
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")
import cv2
import numpy as np
from utils.crop_utils import square_crop_from_bbox

from future_urban_scene_generation_b200 import synth
from oracle import frame_oracle as FO


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


SIZE_PAIRS = [((256, 256), (256, 256)), ((256, 256), (128, 128)), ((256, 256), (97, 97)), ((256, 256), (311, 311)),
              ((256, 256), (412, 412)), ((256, 256), (33, 33)), ((256, 256), (700, 700)), ((256, 256), (255, 255)),
              ((256, 256), (257, 257)), ((64, 48), (256, 256)), ((300, 180), (256, 256)), ((512, 512), (256, 256)),
              ((2, 2), (9, 7)), ((5, 3), (3, 5)), ((1, 7), (4, 4)), ((123, 457), (61, 228)), ((123, 457), (62, 229)),
              ((400, 400), (200, 200)), ((37, 91), (91, 37)), ((256, 256), (1, 1))]
resize_cases = []
for k, ((sh, sw), (dh, dw)) in enumerate(SIZE_PAIRS):
    src = np.random.default_rng(9000 + k).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    ref = cv2.resize(src, (dw, dh))
    got = FO.resize_linear_u8(src, (dw, dh))
    assert np.array_equal(ref, got), ("oracle != cv2", sh, sw, dh, dw)
    resize_cases.append({"seed": 9000 + k, "src_hw": [sh, sw], "dst_hw": [dh, dw], "sha1": sha(ref)})

FRAME_HW = (360, 640)
paste_cases, info_cases = [], []
frame = np.random.default_rng(123).integers(0, 256, FRAME_HW + (3,), dtype=np.uint8)
img_cv, img_or = frame.copy(), frame.copy()
for idx in range(24):
    bbox, mask, net = synth.make_paste_case(idx, FRAME_HW)
    crop, xy_min, pad_b, pad_a, _, _ = square_crop_from_bbox(np.zeros(FRAME_HW + (3,), np.uint8), bbox)
    info = {"crop_xy_min": [int(v) for v in xy_min], "pad_xy_before": [int(v) for v in pad_b], "pad_xy_after": [int(v) for v in pad_a],
            "crop_size_orig": [int(v) for v in crop.shape[:2]]}
    mine = FO.square_crop_info(FRAME_HW, bbox)
    assert all(list(mine[k]) == info[k] for k in info), (idx, mine, info)
    info_cases.append({"idx": idx, "bbox": bbox, **info})
    FO.paste_back(img_cv, net, info, mask, resize=lambda a, ds: cv2.resize(a, ds))
    FO.paste_back(img_or, net, info, mask)
    assert np.array_equal(img_cv, img_or), idx
    paste_cases.append({"idx": idx, "sha1_after": sha(img_cv)})

# ---- VUNet input packing: the reference lines (trajectory_inference.py:205-227) with the reference's own helpers
import torch
import torch.nn.functional as F
from utils.misc_utils import to_tensor
pack_cases = []
for idx in range(16):
    src_sketch_mask, src_sketch_normal, dst_sketch_normal = synth.make_pack_case(idx, FRAME_HW)
    smb = np.bitwise_not(src_sketch_mask)[..., np.newaxis] * frame
    ys, xs = np.nonzero(np.logical_not(src_sketch_mask))
    bb = [np.min(xs), np.min(ys), np.max(xs), np.max(ys)]
    smb = square_crop_from_bbox(smb, bb)[0]
    snb = square_crop_from_bbox(src_sketch_normal, bb)[0]
    dnb = square_crop_from_bbox(dst_sketch_normal, bb)[0]
    smb, snb, dnb = cv2.resize(smb, (256, 256)), cv2.resize(snb, (256, 256)), cv2.resize(dnb, (256, 256))
    smb[np.all(snb == 0, axis=-1)] = 255
    x_1 = F.interpolate(to_tensor(smb).unsqueeze(0), 256)
    x_2 = F.interpolate(to_tensor(snb[..., ::-1].copy()).unsqueeze(0), 256)
    x = torch.cat([x_1, x_2], 1)[0].numpy()
    y = to_tensor(dnb[..., ::-1].copy()).numpy()
    ox, oy, obb = FO.pack_vunet_inputs(frame, src_sketch_mask, src_sketch_normal, dst_sketch_normal)
    assert np.array_equal(ox.view(np.uint32), x.view(np.uint32)) and np.array_equal(oy.view(np.uint32), y.view(np.uint32)), idx
    pack_cases.append({"idx": idx, "bbox": [int(v) for v in bb], "sha1_x": sha(x.astype(np.float32)), "sha1_y": sha(y.astype(np.float32))})

out = {"cv2": cv2.__version__, "pack": pack_cases, "resize": resize_cases, "frame_hw": list(FRAME_HW), "frame_seed": 123, "crop_info": info_cases,
       "paste": paste_cases}
path = os.path.join(ROOT, "tests", "golden", "frame_golden.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, len(resize_cases), "resize cases,", len(paste_cases), "paste steps,", len(pack_cases), "pack cases; oracle == cv2/reference on all")
