"""ctypes front-end of oracle/warp_oracle.c -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's warp half (warp_learn/online_visibility.py:28-150,
warp_learn/planes_utils.py:11-82) plus the OpenCV routines it calls.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under future_urban_scene_generation_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liborc.so")

KP_NAMES = ['left_back_trunk', 'left_back_wheel', 'left_front_light',
            'left_front_wheel', 'right_back_trunk', 'right_back_wheel',
            'right_front_light', 'right_front_wheel', 'upper_left_rearwindow',
            'upper_left_windshield', 'upper_right_rearwindow',
            'upper_right_windshield']          # utils/keypoint_utils.py:9-13
PLANE_NAMES = ['left', 'right', 'roof', 'front', 'back', 'front_bt', 'back_bt']


def build(force=False):
    src = os.path.join(_HERE, "warp_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def plane_table(p):
    idx = (C.c_int * 6)()
    n = lib().orc_plane_table(p, idx)
    return list(idx[:n])


def fill_poly(mask, pts, val=1):
    """cv2.fillPoly(mask, [pts], val) on a single-channel (H,W) uint8 canvas, in place."""
    assert mask.dtype == np.uint8 and mask.ndim == 2 and mask.flags.c_contiguous
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    lib().orc_fill_poly(_p(mask, C.c_uint8), mask.shape[0], mask.shape[1], _p(pts, C.c_int32), len(pts), C.c_uint8(val))
    return mask


def jacobi(A):
    A = np.array(A, np.float64, order="C")
    n = A.shape[0]
    W = np.zeros(n)
    V = np.zeros((n, n))
    lib().orc_jacobi(_p(A, C.c_double), _p(W, C.c_double), _p(V, C.c_double), n)
    return W, V


def solve_eig(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64).ravel()
    x = np.zeros_like(b)
    lib().orc_solve_eig(_p(A, C.c_double), _p(b, C.c_double), _p(x, C.c_double), A.shape[0])
    return x


def invert_eig(A):
    A = np.ascontiguousarray(A, np.float64)
    out = np.zeros_like(A)
    lib().orc_invert_eig(_p(A, C.c_double), _p(out, C.c_double), A.shape[0])
    return out


def find_homography(src, dst, refine=True):
    src = np.ascontiguousarray(src, np.int32).reshape(-1, 2)
    dst = np.ascontiguousarray(dst, np.int32).reshape(-1, 2)
    H = np.zeros(9)
    ok = lib().orc_find_homography(_p(src, C.c_int32), _p(dst, C.c_int32), len(src), _p(H, C.c_double), int(refine))
    return H.reshape(3, 3) if ok else None


def invert3(H):
    H = np.ascontiguousarray(H, np.float64)
    out = np.zeros((3, 3))
    lib().orc_invert3(_p(H, C.c_double), _p(out, C.c_double))
    return out


def warp_perspective(src, H, tapmask=None):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape[:2]
    assert src.shape == (h, w, 3)
    Hm = np.ascontiguousarray(H, np.float64)
    out = np.empty_like(src)
    tm = None
    if tapmask is not None:
        tapmask = np.ascontiguousarray(tapmask, np.uint8)
        tm = _p(tapmask, C.c_uint8)
    lib().orc_warp_perspective(_p(src, C.c_uint8), tm, h, w, _p(Hm, C.c_double), _p(out, C.c_uint8))
    return out


def _E34(E):
    E = np.asarray(E, np.float64)
    if E.shape == (4, 4):
        if not np.all(E[-1] == np.array([0, 0, 0, 1.0])):
            raise ValueError('Format for extrinsic not valid')    # online_visibility.py:47-49
        E = E[:3]
    assert E.shape == (3, 4)
    return np.ascontiguousarray(E)


def kp3d_array(kp3d):
    if isinstance(kp3d, dict):
        kp3d = np.stack([np.asarray(kp3d[k], np.float64) for k in KP_NAMES])
    return np.ascontiguousarray(kp3d, np.float64).reshape(12, 3)


def project_points(K, E, kp3d):
    K = np.ascontiguousarray(K, np.float64)
    E = _E34(E)
    X = np.ascontiguousarray(kp3d, np.float64).reshape(-1, 3)
    uv = np.zeros((len(X), 2))
    lib().orc_project_points(_p(K, C.c_double), _p(E, C.c_double), _p(X, C.c_double), len(X), _p(uv, C.c_double))
    return uv


def plane_distances(E, kp3d):
    E = _E34(E)
    X = kp3d_array(kp3d)
    d = np.zeros(7)
    lib().orc_plane_distances(_p(E, C.c_double), _p(X, C.c_double), _p(d, C.c_double))
    return d


def visibility_from_pts(pts, dist, h, w, return_areas=False):
    pts = np.ascontiguousarray(pts, np.int32).reshape(12, 2)
    dist = np.ascontiguousarray(dist, np.float64)
    vis = np.zeros(7, np.uint8)
    areas = np.zeros(14, np.int32)
    lib().orc_visibility_from_pts(_p(pts, C.c_int32), _p(dist, C.c_double), h, w, _p(vis, C.c_uint8), _p(areas, C.c_int32))
    return (vis, areas.reshape(7, 2)) if return_areas else vis


def compute_visibility(extrinsic, intrinsic, kpoints_3d, h, w, return_pts=False):
    """online_visibility.py:105-150 -> dict of 7 bools (insertion order of PLANE_NAMES)."""
    E = _E34(extrinsic)
    K = np.ascontiguousarray(intrinsic, np.float64)
    X = kp3d_array(kpoints_3d)
    vis = np.zeros(7, np.uint8)
    pts = np.zeros((12, 2), np.int32)
    lib().orc_compute_visibility(_p(E, C.c_double), _p(K, C.c_double), _p(X, C.c_double), h, w, _p(vis, C.c_uint8), _p(pts, C.c_int32))
    d = {n: bool(v) for n, v in zip(PLANE_NAMES, vis)}
    return (d, pts) if return_pts else d


def kp2d_int(kp2d_norm, h, w):
    """planes_utils.py:22-27: normalised [x,y] (dict or (12,2)) -> int32 pixel vertices."""
    if isinstance(kp2d_norm, dict):
        kp2d_norm = np.stack([np.asarray(list(map(float, kp2d_norm[k]))) for k in KP_NAMES])
    a = np.array(kp2d_norm, np.float64).reshape(12, 2).copy()
    a[:, 0] *= w
    a[:, 1] *= h
    return np.int32(a)


def get_planes(image, kp_int):
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape[:2]
    kp = np.ascontiguousarray(kp_int, np.int32).reshape(12, 2)
    planes = np.empty((5, h, w, 3), np.uint8)
    lib().orc_get_planes(_p(image, C.c_uint8), h, w, _p(kp, C.c_int32), _p(planes, C.c_uint8))
    return planes


def warp_unwarp_planes(src_planes, src_kp, dst_kp, src_vis, dst_vis):
    src_planes = np.ascontiguousarray(src_planes, np.uint8)
    _, h, w, _ = src_planes.shape
    skp = np.ascontiguousarray(src_kp, np.int32).reshape(12, 2)
    dkp = np.ascontiguousarray(dst_kp, np.int32).reshape(12, 2)
    sv = np.ascontiguousarray(src_vis, np.uint8)
    dv = np.ascontiguousarray(dst_vis, np.uint8)
    warped = np.empty_like(src_planes)
    unwarped = np.empty_like(src_planes)
    pj = np.zeros(5, np.int8)
    H12 = np.zeros((5, 3, 3))
    lib().orc_warp_unwarp_planes(_p(src_planes, C.c_uint8), h, w, _p(skp, C.c_int32), _p(dkp, C.c_int32),
                                 _p(sv, C.c_uint8), _p(dv, C.c_uint8), _p(warped, C.c_uint8), _p(unwarped, C.c_uint8),
                                 _p(pj, C.c_int8), _p(H12, C.c_double))
    return warped, unwarped, pj, H12


def warp_fused(src, src_kp, dst_kp, K, E_src, E_dst, kp3d):
    """One batch item of the fused path -> (warped (5,H,W,3), vis (2,7), plane_j (5,), H12 (5,3,3))."""
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape[:2]
    skp = np.ascontiguousarray(src_kp, np.int32).reshape(12, 2)
    dkp = np.ascontiguousarray(dst_kp, np.int32).reshape(12, 2)
    K = np.ascontiguousarray(K, np.float64)
    Es, Ed = _E34(E_src), _E34(E_dst)
    X = kp3d_array(kp3d)
    warped = np.empty((5, h, w, 3), np.uint8)
    vis = np.zeros((2, 7), np.uint8)
    pj = np.zeros(5, np.int8)
    H12 = np.zeros((5, 3, 3))
    rc = lib().orc_warp_fused(_p(src, C.c_uint8), h, w, _p(skp, C.c_int32), _p(dkp, C.c_int32), _p(K, C.c_double),
                              _p(Es, C.c_double), _p(Ed, C.c_double), _p(X, C.c_double), _p(warped, C.c_uint8),
                              _p(vis, C.c_uint8), _p(pj, C.c_int8), _p(H12, C.c_double))
    if rc != 0:
        raise ValueError(f"orc_warp_fused rc={rc}")
    return warped, vis, pj, H12
