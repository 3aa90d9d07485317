"""Per-step keypoint kinematics of the trajectory loop, batched (include/fusg.h:
fusg_step_keypoints).

`traj_test` moves the 12 CAD keypoints of every vehicle along its predicted path, re-projects them with
cv2.projectPoints and truncates them into plane vertices, once per future step inside a Python loop
(trajectory_inference.py:267-298, :359-367; warp_learn/vehicle_utils.py:24-26; warp_learn/planes_utils.py:22-27).  Here the
few scalars per step (heading, distance, the +-20 degree gates) are computed on the host exactly as the reference does, and
everything per keypoint runs in one launch for all (vehicle, step) pairs, producing the `kp3d_dst` / `dst_kp` arrays
`warp_batch(..., kp3d_dst=...)` takes.
"""
import numpy as np

from . import _lib


def z_rot(alpha):
    """utils/geometry.py:80-113, counter-clockwise numpy form: a float32 matrix."""
    cz, sz = np.cos(alpha), np.sin(alpha)
    return np.asarray([[cz, -sz, 0.], [sz, cz, 0.], [0., 0., 1.]], dtype=np.float32)


def trajectory_poses(meter_coords):
    """Heading and translation of every future position of one vehicle (trajectory_inference.py:258-298).
    meter_coords (S+1, 2): row 0 is the current position.  Returns theta (S,) f64, tr (S,3) f64, rot (S,3,3) f32 with
    rot[s] = z_rot(theta[s]) -- the keypoints always rotate by theta (:361); `tr` uses z_rot(0) when a gate trips."""
    mc = np.asarray(meter_coords, dtype=np.float64)
    if mc.ndim != 2 or mc.shape[1] != 2 or len(mc) < 2:
        raise ValueError("meter_coords must be (S+1, 2) with S >= 1")
    x_start, y_start = mc[0]
    theta_start = np.arctan2(np.mean(mc[1:20, 1] - y_start), np.mean(mc[1:20, 0] - x_start))
    S = len(mc) - 1
    theta = np.empty(S)
    tr = np.empty((S, 3))
    rot = np.empty((S, 3, 3), dtype=np.float32)
    for n in range(1, S + 1):
        x_cur, y_cur = mc[n]
        distance = np.linalg.norm(mc[0] - mc[n])
        th = np.arctan2(y_cur - y_start, x_cur - x_start) - theta_start
        if 1 < n < S - 1:
            cur_theta = np.degrees(np.arctan2(y_cur - mc[n - 1, 1], x_cur - mc[n - 1, 0]))
            next_theta = np.degrees(np.arctan2(mc[n + 1, 1] - y_cur, mc[n + 1, 0] - x_cur))
            gate = -20 < cur_theta - next_theta < 20
        else:
            gate = -20 < np.degrees(th) < 20
        delta_t = np.zeros(3)
        delta_t[1] = -distance
        theta[n - 1] = th
        tr[n - 1] = delta_t @ z_rot(th if gate else 0)       # a single product per component: no summation-order freedom
        rot[n - 1] = z_rot(th)
    return theta, tr, rot


def step_keypoints_batch(kp3d, vehicle, rot, tr, R, t, K, h, w, device=None):
    """All (vehicle, step) items at once.
    kp3d (V,12,3) f64 CAD keypoints (_KP_NAMES order); vehicle (N,) int index of each item's vehicle; rot (N,3,3) f32 and
    tr (N,3) f64 from `trajectory_poses`; R (V,3,3) f64 = cv2.Rodrigues(rvect)[0], t (V,3), K (V,3,3) or (3,3); frame h, w.
    Returns device tensors (kp3d_dst (N,12,3) f64, kp2d (N,12,2) f64, dst_kp (N,12,2) i32)."""
    torch = _lib.require_cuda()
    device = torch.device(device if device is not None else "cuda")

    def dev(a, dtype, shape):
        x = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a)
        return x.to(dtype).reshape(shape).to(device).contiguous()

    V = int(np.asarray(kp3d.shape if hasattr(kp3d, "shape") else np.shape(kp3d))[0])
    N = int(len(vehicle))
    Kt = torch.as_tensor(K)
    if Kt.dim() == 2:
        Kt = Kt.unsqueeze(0).expand(V, 3, 3)
    X, veh = dev(kp3d, torch.float64, (V, 12, 3)), dev(vehicle, torch.int32, (N,))
    if N and (int(veh.min()) < 0 or int(veh.max()) >= V):
        raise ValueError("vehicle index out of range")
    rot_d, tr_d = dev(rot, torch.float32, (N, 3, 3)), dev(tr, torch.float64, (N, 3))
    R_d, t_d, K_d = dev(R, torch.float64, (V, 3, 3)), dev(t, torch.float64, (V, 3)), dev(Kt, torch.float64, (V, 3, 3))
    moved = torch.empty((N, 12, 3), dtype=torch.float64, device=device)
    kp2d = torch.empty((N, 12, 2), dtype=torch.float64, device=device)
    verts = torch.empty((N, 12, 2), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        rc = _lib.lib().fusg_step_keypoints(_lib.ptr(X), _lib.ptr(veh), _lib.ptr(rot_d), _lib.ptr(tr_d), _lib.ptr(R_d), _lib.ptr(t_d), _lib.ptr(K_d),
                                            _lib.ptr(moved), _lib.ptr(kp2d), _lib.ptr(verts), N, int(h), int(w), _lib.stream_ptr(torch))
    _lib.check(rc, "fusg_step_keypoints")
    moved._keep = (X, veh, rot_d, tr_d, R_d, t_d, K_d)
    return moved, kp2d, verts
