"""CPU restatement of the per-step keypoint kinematics of the trajectory loop -- TEST INFRASTRUCTURE ONLY.

Follows the reference lines:
  trajectory_inference.py:258-298   heading / distance of every future position, the +-20 degree gates, `tr`
  trajectory_inference.py:359-367   `v @ z_rot(theta) + tr` for the 12 CAD keypoints, cv2.projectPoints
  utils/geometry.py:80-113          z_rot (a float32 matrix)          utils/geometry.py:140-144  get_delta_t_vec
  warp_learn/vehicle_utils.py:24-26 normalize_kpoints (/ w, / h)      warp_learn/planes_utils.py:22-27  * w, * h, np.int32

Third-party arithmetic restated here (not vendored under /root/reference; pinned by scripts/make_golden_kinematics.py against
the wheels of the build container, numpy 2.3.5 + OpenBLAS and opencv-python 4.13.0):
  * `v @ M` for a float64 (3,) vector and the float32 (3,3) z_rot matrix: numpy promotes M to float64 and its BLAS evaluates
    every output as the FMA chain fma(v2, M2c, fma(v1, M1c, v0 * M0c)) (20,000 random cases, 0 mismatches; the plain
    left-to-right sum differs in 40 % of them).  Emulated exactly with rationals.
  * cv2.projectPoints with zero distortion: X_cam = R X + t evaluated left to right, z -> 1/z, x*fx + cx (the distortion
    polynomial is exactly 1 and the tilt matrix the identity) -- bit-exact on 3,000 random cases; R = cv2.Rodrigues(rvec)
    is taken from OpenCV once per vehicle (the reference computes it the same way, utils/geometry.py:216).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from fractions import Fraction

import numpy as np


def fma(a, b, c):
    """Correctly rounded a*b + c in float64."""
    a, b, c = float(a), float(b), float(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return a * b + c
    r = Fraction(a) * Fraction(b) + Fraction(c)
    if r == 0:
        return a * b + c           # keeps IEEE signed-zero behaviour
    return float(r)


def z_rot_f32(theta):
    """utils/geometry.py:80-113 (counter-clockwise): a float32 matrix built from float64 cos / sin."""
    cz, sz = np.cos(theta), np.sin(theta)
    return np.asarray([[cz, -sz, 0.], [sz, cz, 0.], [0., 0., 1.]], dtype=np.float32)


def vec_mat(v, M):
    """numpy's float64 (3,) @ float32 (3,3) as the build container evaluates it (see the module docstring)."""
    M = np.asarray(M, dtype=np.float64)
    return np.array([fma(v[2], M[2, c], fma(v[1], M[1, c], float(v[0]) * float(M[0, c]))) for c in range(3)], dtype=np.float64)


def trajectory_poses(meter_coords):
    """trajectory_inference.py:258-298 for one vehicle: meter_coords (S+1, 2) -> theta (S,), tr (S,3), rot (S,3,3) float32.
    The keypoints always rotate by theta (:361); only the translation direction is gated (:290-298)."""
    mc = np.asarray(meter_coords, dtype=np.float64)
    x_start, y_start = mc[0]
    delta_x = np.mean(mc[1:20, 0] - x_start)
    delta_y = np.mean(mc[1:20, 1] - y_start)
    theta_start = np.arctan2(delta_y, delta_x)
    n_future = len(mc) - 1
    thetas, trs, rots = [], [], []
    for n in range(1, n_future + 1):
        cur = mc[n]
        distance = np.linalg.norm(mc[0] - cur)
        theta = np.arctan2(cur[1] - y_start, cur[0] - x_start) - theta_start
        delta_t = np.zeros(3)
        delta_t[1] = -distance
        if 1 < n < n_future - 1:
            cur_theta = np.degrees(np.arctan2(cur[1] - mc[n - 1, 1], cur[0] - mc[n - 1, 0]))
            next_theta = np.degrees(np.arctan2(mc[n + 1, 1] - cur[1], mc[n + 1, 0] - cur[0]))
            gate = -20 < cur_theta - next_theta < 20
        else:
            gate = -20 < np.degrees(theta) < 20
        tr = vec_mat(delta_t, z_rot_f32(theta if gate else 0))
        thetas.append(theta)
        trs.append(tr)
        rots.append(z_rot_f32(theta))
    return np.asarray(thetas), np.asarray(trs), np.asarray(rots)


def step_keypoints(kp3d, rot_f32, tr):
    """trajectory_inference.py:359-361: every keypoint v -> v @ z_rot(theta) + tr.   kp3d (12,3) f64."""
    return np.stack([vec_mat(v, rot_f32) + np.asarray(tr, dtype=np.float64) for v in np.asarray(kp3d, dtype=np.float64)])


def project_points_cv(P, R, t, K):
    """cv2.projectPoints(P, rvec, tvec, K, zeros) with R = cv2.Rodrigues(rvec)[0] (trajectory_inference.py:363-367)."""
    R, t, K = np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3), np.asarray(K, np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    out = np.empty((len(P), 2), dtype=np.float64)
    for i, (X, Y, Z) in enumerate(np.asarray(P, dtype=np.float64)):
        x = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + t[0]
        y = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + t[1]
        z = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + t[2]
        z = 1. / z if z else 1.
        x *= z
        y *= z
        out[i] = (x * fx + cx, y * fy + cy)
    return out


def plane_vertices(kp2d, w, h):
    """vehicle_utils.py:24-26 + planes_utils.py:22-27: (x / w) * w, (y / h) * h, then np.int32 (truncation)."""
    kp = np.array(kp2d, dtype=np.float64)
    kp[:, 0] /= w
    kp[:, 1] /= h
    kp[:, 0] *= w
    kp[:, 1] *= h
    return np.int32(kp)


def step(kp3d, rot_f32, tr, R, t, K, w, h):
    """One (vehicle, future step): moved 3D keypoints, their cv2 projection, the int32 plane vertices."""
    moved = step_keypoints(kp3d, rot_f32, tr)
    kp2d = project_points_cv(moved, R, t, K)
    return moved, kp2d, plane_vertices(kp2d, w, h)
