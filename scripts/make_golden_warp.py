"""Runs the imported reference (cv2 + numpy, /root/reference) on seeded synthetic crops and commits
what the oracle must reproduce: tests/golden/warp_golden.json

  per crop (160 detailed cases): visibility dicts of both poses, get_planes sha1 per plane, plane
            vertices, warp_unwarp_planes()[0] sha1 per plane;
  bulk (seeds 1000..2999): one sha1 per crop over (visibility of both poses, warped planes) of the
            REFERENCE's outputs -- the oracle and the CUDA path must reproduce every one of them.
The generator asserts that the oracle is bit-identical to the reference on every plane it writes.
Only runs where /root/reference exists (the build container)."""
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
from warp_learn.online_visibility import compute_visibility, pascal_texture_planes
from warp_learn.planes_utils import get_planes, warp_unwarp_planes

from oracle import warp_oracle as O
from future_urban_scene_generation_b200 import synth


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


N = 160
H = W = 256
cases = []
n_planes = n_equal = 0
for idx in range(N):
    p = synth.make_pose_pair(idx)
    img = synth.make_crop(idx)
    kp3d = {k: p["kp3d"][i] for i, k in enumerate(synth.KP_NAMES)}
    vs = compute_visibility(p["E_src"], p["K"], kp3d, H, W)
    vd = compute_visibility(p["E_dst"], p["K"], kp3d, H, W)
    ks = {k: p["kp2d_src"][i] for i, k in enumerate(synth.KP_NAMES)}
    kd = {k: p["kp2d_dst"][i] for i, k in enumerate(synth.KP_NAMES)}
    sp, skp, sv = get_planes(img, ks, 'car', vs)
    dp, dkp, dv = get_planes(img, kd, 'car', vd)
    wr, _ = warp_unwarp_planes(sp, skp, dkp, sv, dv, 'car', pascal_texture_planes)
    wo, vis, pj, H12 = O.warp_fused(img, p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
    assert [bool(vs[k]) for k in O.PLANE_NAMES] == [bool(v) for v in vis[0]], idx
    assert [bool(vd[k]) for k in O.PLANE_NAMES] == [bool(v) for v in vis[1]], idx
    assert np.array_equal(sp, O.get_planes(img, p["src_kp"])), idx
    assert np.array_equal(wr, wo), idx
    written = [bool(wr[j].any()) for j in range(5)]
    n_planes += sum(written)
    cases.append({
        "idx": idx,
        "vis_src": [int(bool(vs[k])) for k in O.PLANE_NAMES], "vis_dst": [int(bool(vd[k])) for k in O.PLANE_NAMES],
        "src_kp": p["src_kp"].tolist(), "dst_kp": p["dst_kp"].tolist(),
        "planes_sha1": [sha(sp[j]) for j in range(5)],
        "warped_sha1": [sha(wr[j]) for j in range(5)],
        "warped_nonzero": written,
        "plane_j": [int(v) for v in pj],
    })

def ref_crop(idx, H=H, W=W, out_of_frame=False):
    p = synth.make_pose_pair(idx, H, W, out_of_frame=out_of_frame)
    img = synth.make_crop(idx, H, W)
    kp3d = {k: p["kp3d"][i] for i, k in enumerate(synth.KP_NAMES)}
    vs = compute_visibility(p["E_src"], p["K"], kp3d, H, W)
    vd = compute_visibility(p["E_dst"], p["K"], kp3d, H, W)
    ks = {k: p["kp2d_src"][i] for i, k in enumerate(synth.KP_NAMES)}
    kd = {k: p["kp2d_dst"][i] for i, k in enumerate(synth.KP_NAMES)}
    sp, skp, sv = get_planes(img, ks, 'car', vs)
    dp, dkp, dv = get_planes(img, kd, 'car', vd)
    wr, _ = warp_unwarp_planes(sp, skp, dkp, sv, dv, 'car', pascal_texture_planes)
    vis = np.array([[int(bool(v[k])) for k in O.PLANE_NAMES] for v in (vs, vd)], np.uint8)
    wo, vo, pj, _ = O.warp_fused(img, p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
    assert np.array_equal(vis, vo[:2]) and np.array_equal(wr, wo) and np.array_equal(sp, O.get_planes(img, p["src_kp"])), (idx, H, W)
    return vis, wr


def bulk_set(first, n, H=H, W=W, out_of_frame=False):
    shas, planes = [], 0
    for idx in range(first, first + n):
        vis, wr = ref_crop(idx, H, W, out_of_frame)
        planes += sum(bool(wr[j].any()) for j in range(5))
        shas.append(hashlib.sha1(vis.tobytes() + np.ascontiguousarray(wr).tobytes()).hexdigest()[:16])
    return shas, planes


import warnings
warnings.simplefilter("ignore")          # normalize_kpoints warns about keypoints > 1.0 in the out-of-frame sets
BULK0, BULKN = 1000, 2000
bulk, bulk_planes = bulk_set(BULK0, BULKN)
# vehicles leaving the frame (clipped-polygon regime of cv2.fillPoly): 2000 crops at 256 x 256, 60 at the reference's
# own 1280 x 720 working resolution (GUI/app_interface.py:181)
oob, oob_planes = bulk_set(0, 2000, out_of_frame=True)
oob720, oob720_planes = bulk_set(0, 60, 720, 1280, out_of_frame=True)

gold = {"cv2_version": cv2.__version__, "n": N, "hw": [H, W], "written_planes": n_planes, "cases": cases,
        "bulk_first": BULK0, "bulk_sha1_16": bulk, "bulk_written_planes": bulk_planes,
        "oob_sha1_16": oob, "oob_written_planes": oob_planes,
        "oob720_sha1_16": oob720, "oob720_written_planes": oob720_planes}
json.dump(gold, open(os.path.join(ROOT, "tests", "golden", "warp_golden.json"), "w"))
print(f"{N} crops: {n_planes} written planes, all identical to the oracle; bulk {BULKN} crops, {bulk_planes} written planes; "
      f"out-of-frame 2000 crops / {oob_planes} planes at 256x256, 60 crops / {oob720_planes} planes at 1280x720")
