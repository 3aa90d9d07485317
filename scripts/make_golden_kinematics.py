"""Pins oracle/kinematics_oracle.py against the reference lines of the trajectory loop executed with the reference's own
helpers (utils/geometry.py z_rot / get_delta_t_vec, utils/keypoint_utils.py normalize_kpoints / kpoints_*), numpy and
cv2.projectPoints in the build container, and writes tests/golden/kinematics_golden.json (bit patterns as SHA-256 plus the
OpenCV-derived rotation matrices, so the GPU box needs neither the reference nor cv2)."""
import hashlib
import json
import os
import sys
import types

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Stub(types.ModuleType):                  # open3d / matplotlib are absent; the helpers used here never touch them
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(k)


for name in ("open3d", "matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
    sys.modules[name] = _Stub(name)
sys.path.append("/root/reference")
import warnings
warnings.filterwarnings("ignore")
import cv2
import numpy as np

from utils.geometry import z_rot, get_delta_t_vec                                    # the reference
from utils.keypoint_utils import kpoints_array_to_dict, kpoints_dict_to_array, normalize_kpoints, _KP_NAMES
from oracle import kinematics_oracle as KO
from future_urban_scene_generation_b200 import synth

assert list(_KP_NAMES) == synth.KP_NAMES
pascal_car = {'left': ['left_front_wheel', 'left_back_wheel', 'left_back_trunk', 'upper_left_rearwindow', 'upper_left_windshield', 'left_front_light']}


def reference_steps(case, rvect, tvect):
    """trajectory_inference.py:258-298 and :359-367, vehicle_utils.py:24-26, planes_utils.py:22-27 -- verbatim control flow."""
    meter_coords, K = case["meter_coords"], case["K"]
    dist = np.zeros((1, 5), dtype=np.float32)
    orig_kpoints_3d_dict = kpoints_array_to_dict(case["kp3d"].copy())
    x_start, y_start = meter_coords[0]
    delta_x = np.mean(meter_coords[1:20, 0] - x_start)
    delta_y = np.mean(meter_coords[1:20, 1] - y_start)
    theta_start = np.arctan2(delta_y, delta_x)
    out = []
    for n, cur_pos in enumerate(meter_coords[1:], 1):
        distance = np.linalg.norm(meter_coords[0] - cur_pos)
        x_cur, y_cur = cur_pos
        theta = np.arctan2(y_cur - y_start, x_cur - x_start) - theta_start
        delta_t = get_delta_t_vec('y', -distance)
        if 1 < n < len(meter_coords[1:]) - 1:
            cur_theta = np.degrees(np.arctan2(y_cur - meter_coords[n - 1, 1], x_cur - meter_coords[n - 1, 0]))
            next_theta = np.degrees(np.arctan2(meter_coords[n + 1, 1] - y_cur, meter_coords[n + 1, 0] - x_cur))
            theta_diff = cur_theta - next_theta
            tr = delta_t @ z_rot(theta) if -20 < theta_diff < 20 else delta_t @ z_rot(0)
        else:
            tr = delta_t @ z_rot(theta) if -20 < np.degrees(theta) < 20 else delta_t @ z_rot(0)
        kpoints_3d_dict = orig_kpoints_3d_dict.copy()
        for k, v in kpoints_3d_dict.items():
            kpoints_3d_dict[k] = v @ z_rot(theta) + tr
        moved = kpoints_dict_to_array(kpoints_3d_dict, dim=3)
        kpoints_2d_next, _ = cv2.projectPoints(moved, rvect, tvect, K, dist)
        kpoints_2d_next = kpoints_2d_next.squeeze(1)
        kp2d = kpoints_2d_next.copy()
        norm = normalize_kpoints(kpoints_dict_to_array(kpoints_array_to_dict(kpoints_2d_next)), max_x=case["w"], max_y=case["h"])
        d = kpoints_array_to_dict(norm)
        p = np.asarray([list(map(float, d[k])) for k in _KP_NAMES])
        p[:, 0] *= case["w"]
        p[:, 1] *= case["h"]
        out.append((theta, tr, moved, kp2d, np.int32(p)))
    return out


gold = {"numpy": np.__version__, "cv2": cv2.__version__, "cases": []}
n_items = 0
for idx in range(9):
    case = synth.make_trajectory_case(idx)
    rvect, _ = cv2.Rodrigues(case["R"])
    R_cv, _ = cv2.Rodrigues(rvect)                       # what projectPoints uses internally
    tvect = case["t"].reshape(3, 1)
    ref = reference_steps(case, rvect, tvect)
    thetas, trs, rots = KO.trajectory_poses(case["meter_coords"])
    hm, h2, hv = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    for s, (theta, tr, moved, kp2d, verts) in enumerate(ref):
        assert theta == thetas[s] and np.array_equal(tr, trs[s]) and np.array_equal(z_rot(theta), rots[s]), (idx, s)
        o_moved, o_kp2d, o_verts = KO.step(case["kp3d"], rots[s], trs[s], R_cv, case["t"], case["K"], case["w"], case["h"])
        assert np.array_equal(moved, o_moved), (idx, s, "moved")
        assert np.array_equal(kp2d, o_kp2d), (idx, s, "kp2d")
        assert np.array_equal(verts, o_verts), (idx, s, "verts")
        hm.update(moved.tobytes()); h2.update(kp2d.tobytes()); hv.update(verts.tobytes())
        n_items += 1
    ungated = sum(1 for s in range(len(ref)) if not np.array_equal(trs[s], KO.vec_mat(np.array([0.0, -np.linalg.norm(case["meter_coords"][0] - case["meter_coords"][s + 1]), 0.0]), rots[s])))
    gold["cases"].append({"idx": idx, "R_cv_hex": [float(v).hex() for v in R_cv.flatten()], "steps": len(ref), "translation_gated_steps": ungated,
                          "sha256_moved": hm.hexdigest(), "sha256_kp2d": h2.hexdigest(), "sha256_verts": hv.hexdigest(),
                          "first_verts": ref[0][4].tolist(), "last_kp2d_hex": [float(v).hex() for v in ref[-1][3].flatten()]})
    print(idx, "steps", len(ref), "translation gated off in", ungated, "steps; bit-identical to the reference lines")
with open(os.path.join(ROOT, "tests", "golden", "kinematics_golden.json"), "w") as f:
    json.dump(gold, f, indent=1)
print("wrote tests/golden/kinematics_golden.json:", n_items, "items")
