"""Pins oracle/warp_oracle.c::orc_find_homography (n = 6: normalised DLT + the 9-parameter LM
refinement of the installed OpenCV) against cv2.findHomography, bit for bit, on the side-plane
point sets of seeded synthetic pose pairs -- and the written planes of warp_unwarp_planes()[0]
against the imported reference when /root/reference exists.

    python scripts/check_lm_vs_cv2.py [first_seed] [n_seeds]

Build-container only (needs cv2); the GPU box never runs this."""
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
import numpy as np

from oracle import warp_oracle as O
from future_urban_scene_generation_b200 import synth


def main(first=1000, n=2000):
    bad = tot = none = 0
    worst = 0.0
    for idx in range(first, first + n):
        p = synth.make_pose_pair(idx)
        for pl in (0, 1):
            ids = O.plane_table(pl)
            for s, d in ((p["src_kp"][ids], p["dst_kp"][ids]), (p["dst_kp"][ids], p["src_kp"][ids])):
                Hc, _ = cv2.findHomography(s, d)
                Ho = O.find_homography(s, d)
                tot += 1
                if Hc is None or Ho is None:
                    none += 1
                    bad += (Hc is None) != (Ho is None)
                    continue
                if not np.array_equal(Hc, Ho):
                    bad += 1
                    worst = max(worst, float(np.abs(Hc - Ho).max()))
    print(f"cv2 {cv2.__version__}: {tot} six-point findHomography calls, {none} None, "
          f"{bad} not bit-identical (worst abs diff {worst:.3g})")
    return bad


def random_sets(n=4000, seed=7):
    """Unstructured 6-point sets (not car-shaped): stresses the eigenvalue cut and the lambda == 0 branch.
    The path only ever solves 4- and 6-point planes (online_visibility.py:9-25); for n = 5, 7, 8 about 1 %
    of such sets still differ from cv2 in the 8th digit (other row counts take other summation
    paths inside cv::norm / cv::gemm) -- unpinned and unused."""
    rng = np.random.default_rng(seed)
    bad = 0
    for _ in range(n):
        k = 6
        s = rng.integers(0, 256, (k, 2)).astype(np.int32)
        Ht = np.eye(3) + rng.normal(0, [[0.1, 0.1, 10], [0.1, 0.1, 10], [2e-4, 2e-4, 0]])
        q = np.c_[s, np.ones(k)] @ Ht.T
        d = (q[:, :2] / q[:, 2:] + rng.normal(0, 1.0, (k, 2))).astype(np.int32)
        Hc, _ = cv2.findHomography(s, d)
        Ho = O.find_homography(s, d)
        if (Hc is None) != (Ho is None) or (Hc is not None and not np.array_equal(Hc, Ho)):
            bad += 1
    print(f"random 6-point sets: {n} calls, {bad} not bit-identical")
    return bad


def planes_vs_reference(first=1000, n=2000):
    """warp_unwarp_planes()[0] of the imported reference vs the oracle, every written plane."""
    if not os.path.isdir("/root/reference"):
        print("no /root/reference: plane comparison skipped")
        return 0
    sys.path.append("/root/reference")
    from warp_learn.online_visibility import compute_visibility, pascal_texture_planes
    from warp_learn.planes_utils import get_planes, warp_unwarp_planes
    written = differ = 0
    for idx in range(first, first + n):
        p = synth.make_pose_pair(idx)
        img = synth.make_crop(idx)
        kp3d = {k: p["kp3d"][i] for i, k in enumerate(synth.KP_NAMES)}
        vs = compute_visibility(p["E_src"], p["K"], kp3d, 256, 256)
        vd = compute_visibility(p["E_dst"], p["K"], kp3d, 256, 256)
        ks = {k: p["kp2d_src"][i] for i, k in enumerate(synth.KP_NAMES)}
        kd = {k: p["kp2d_dst"][i] for i, k in enumerate(synth.KP_NAMES)}
        sp, skp, sv = get_planes(img, ks, 'car', vs)
        dp, dkp, dv = get_planes(img, kd, 'car', vd)
        wr, _ = warp_unwarp_planes(sp, skp, dkp, sv, dv, 'car', pascal_texture_planes)
        wo, vis, pj, H12 = O.warp_fused(img, p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
        for j in range(5):
            if wr[j].any() or wo[j].any():
                written += 1
                differ += not np.array_equal(wr[j], wo[j])
    print(f"warp_unwarp_planes()[0] on seeds {first}..{first + n - 1}: {written} written planes, {differ} differ from the reference")
    return differ


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:3]]
    rc = main(*a) + random_sets() + planes_vs_reference(*a)
    sys.exit(1 if rc else 0)
