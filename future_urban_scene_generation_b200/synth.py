"""Deterministic synthetic inputs for the novel-view-completion hot path.

No Pascal3D CAD YAMLs, checkpoints or videos are available offline, so tests, bench.py and
smoke() all draw from this generator: a 5 m box-car with the 12 Pascal3D car keypoints
(utils/keypoint_utils.py:9-13 order), a pinhole camera looking at the origin, and uniform-noise
source crops (the worst case for interpolation parity; bandwidth is data independent).

Everything is numpy on the host -- this module produces inputs, it computes nothing on the path.
"""
import numpy as np

KP_NAMES = ['left_back_trunk', 'left_back_wheel', 'left_front_light',
            'left_front_wheel', 'right_back_trunk', 'right_back_wheel',
            'right_front_light', 'right_front_wheel', 'upper_left_rearwindow',
            'upper_left_windshield', 'upper_right_rearwindow',
            'upper_right_windshield']

# CAD axes: x = left(-)/right(+), y = back(-)/front(+), z = up
_BOX_CAR = {
    'left_back_trunk': (-0.9, -2.5, 0.9), 'right_back_trunk': (0.9, -2.5, 0.9),
    'left_back_wheel': (-0.9, -1.5, 0.3), 'right_back_wheel': (0.9, -1.5, 0.3),
    'left_front_wheel': (-0.9, 1.5, 0.3), 'right_front_wheel': (0.9, 1.5, 0.3),
    'left_front_light': (-0.9, 2.5, 0.7), 'right_front_light': (0.9, 2.5, 0.7),
    'upper_left_rearwindow': (-0.7, -1.2, 1.5), 'upper_right_rearwindow': (0.7, -1.2, 1.5),
    'upper_left_windshield': (-0.7, 0.6, 1.5), 'upper_right_windshield': (0.7, 0.6, 1.5),
}


def cad_keypoints(cad_id: int = 0) -> np.ndarray:
    """(12,3) fp64 keypoints of synthetic CAD `cad_id` (per-id jitter of +-5 %)."""
    base = np.array([_BOX_CAR[k] for k in KP_NAMES], np.float64)
    jit = np.random.default_rng(cad_id).uniform(-0.05, 0.05, base.shape)
    return base * (1.0 + jit)


def intrinsic(h: int = 256, w: int = 256) -> np.ndarray:
    f = 300.0 * w / 256.0
    return np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]], np.float64)


def look_at_extrinsic(azimuth_deg: float, elevation_deg: float, distance: float) -> np.ndarray:
    """World->camera 4x4 for a camera on a sphere around the origin, looking at it, z up."""
    az, el = np.deg2rad(azimuth_deg), np.deg2rad(elevation_deg)
    c = distance * np.array([np.cos(el) * np.sin(az), -np.cos(el) * np.cos(az), np.sin(el)])
    fwd = -c / np.linalg.norm(c)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])           # rows: camera x, y(down), z(forward)
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ c
    return E


def project(K: np.ndarray, E: np.ndarray, X: np.ndarray) -> np.ndarray:
    """The reference's pinhole model (warp_learn/online_visibility.py:28-56), vectorised."""
    Xh = np.concatenate([X, np.ones((len(X), 1))], 1)
    p = K @ E[:3] @ Xh.T
    p = p / p[2]
    return p.T[:, :2]


def make_pose_pair(idx: int, h: int = 256, w: int = 256, max_tries: int = 200, out_of_frame: bool = False):
    """Source/destination poses for crop `idx` with every projected keypoint inside the frame, or -- with
    `out_of_frame` -- a vehicle leaving the frame (trajectory_inference.py:374-379): the look-at point is shifted so
    that the car straddles a border, and at least one keypoint of either pose must project outside.

    Returns dict(K, E_src, E_dst, kp3d, kp2d_src, kp2d_dst, src_kp, dst_kp) where kp2d_* are the
    normalised (12,2) arrays the reference passes to get_planes and *_kp their int32 truncation
    (planes_utils.py:22-27).
    """
    rng = np.random.default_rng((1234 if not out_of_frame else 91_234) + idx)
    K = intrinsic(h, w)
    kp3d = cad_keypoints(idx % 10)
    for _ in range(max_tries):
        az = rng.uniform(0, 360)
        el = rng.uniform(5, 40)
        E_src = look_at_extrinsic(az, el, rng.uniform(8, 11))
        E_dst = look_at_extrinsic(az + rng.uniform(-20, 20), el, rng.uniform(8, 11))
        if out_of_frame:
            # slide the principal point: the car drifts towards / across a border of the frame
            K = intrinsic(h, w)
            K[0, 2] += rng.uniform(-0.6, 0.6) * w
            K[1, 2] += rng.uniform(-0.6, 0.6) * h
        p_src, p_dst = project(K, E_src, kp3d), project(K, E_dst, kp3d)
        inside = True
        for p in (p_src, p_dst):
            if p[:, 0].min() < 0 or p[:, 0].max() > w - 1 or p[:, 1].min() < 0 or p[:, 1].max() > h - 1:
                inside = False
        if inside != out_of_frame:
            break
    else:
        raise RuntimeError("no suitable pose found")
    out = dict(K=K, E_src=E_src, E_dst=E_dst, kp3d=kp3d)
    for name, p in (("src", p_src), ("dst", p_dst)):
        norm = p / np.array([w, h], np.float64)
        out[f"kp2d_{name}"] = norm
        px = norm.copy()
        px[:, 0] *= w
        px[:, 1] *= h
        out[f"{name}_kp"] = np.int32(px)
    return out


def make_crop(idx: int, h: int = 256, w: int = 256) -> np.ndarray:
    """Uniform-noise uint8 source crop (h,w,3) for crop `idx`."""
    return np.random.default_rng(77_000 + idx).integers(0, 256, (h, w, 3), dtype=np.uint8)


def make_warp_batch(start: int, count: int, h: int = 256, w: int = 256, crops: bool = True, out_of_frame: bool = False):
    """Stacked arrays for crops [start, start+count): the layout the C ABI takes (include/fusg.h)."""
    poses = [make_pose_pair(start + i, h, w, out_of_frame=out_of_frame) for i in range(count)]
    batch = dict(
        K=np.stack([p["K"] for p in poses]),
        E_src=np.stack([p["E_src"][:3] for p in poses]),
        E_dst=np.stack([p["E_dst"][:3] for p in poses]),
        kp3d=np.stack([p["kp3d"] for p in poses]),
        src_kp=np.stack([p["src_kp"] for p in poses]),
        dst_kp=np.stack([p["dst_kp"] for p in poses]),
    )
    if crops:
        batch["src"] = np.stack([make_crop(start + i, h, w) for i in range(count)])
    return batch


def make_vunet_inputs(start: int, count: int, res: int = 256):
    """(x (B,6,res,res), y_tilde (B,3,res,res)) fp32 in [-1,1] = to_tensor of uint8 noise
    (utils/misc_utils.py:35-50)."""
    xs, ys = [], []
    for i in range(count):
        rng = np.random.default_rng(55_000 + start + i)
        x8 = rng.integers(0, 256, (6, res, res), dtype=np.uint8)
        y8 = rng.integers(0, 256, (3, res, res), dtype=np.uint8)
        xs.append(np.float32(x8) / 255 * 2.0 - 1.0)
        ys.append(np.float32(y8) / 255 * 2.0 - 1.0)
    return np.stack(xs).astype(np.float32), np.stack(ys).astype(np.float32)


def make_icn_inputs(start: int, count: int, res: int = 256, channels: int = 21):
    """ICN generator input (B,21,res,res) fp32 in [-1,1]: the value grid of `get_icn_inputs`
    (warp_learn/models.py:354-366: Normalize(0.5, 0.5)(ToTensor(uint8))) filled with uint8 noise."""
    xs = []
    for i in range(count):
        rng = np.random.default_rng(77_000 + start + i)
        x8 = rng.integers(0, 256, (channels, res, res), dtype=np.uint8)
        xs.append((np.float32(x8) / 255 - 0.5) / 0.5)
    return np.stack(xs).astype(np.float32)


def make_vunet_inputs_u8(start: int, count: int, res: int = 256):
    """The same data as `make_vunet_inputs`, as the three uint8 images the reference holds before `to_tensor`
    (trajectory_inference.py:215-220): (src_sketch_mask_bbox, src_sketch_normal_bbox, dst_sketch_normal_bbox), each
    (B,res,res,3) -- `to_tensor` + `[..., ::-1]` + `cat` of them is exactly `make_vunet_inputs(start, count)`."""
    m, ns, nd = [], [], []
    for i in range(count):
        rng = np.random.default_rng(55_000 + start + i)
        x8 = rng.integers(0, 256, (6, res, res), dtype=np.uint8)
        y8 = rng.integers(0, 256, (3, res, res), dtype=np.uint8)
        m.append(np.transpose(x8[0:3], (1, 2, 0)))
        ns.append(np.transpose(x8[3:6][::-1], (1, 2, 0)))
        nd.append(np.transpose(y8[::-1], (1, 2, 0)))
    return np.ascontiguousarray(np.stack(m)), np.ascontiguousarray(np.stack(ns)), np.ascontiguousarray(np.stack(nd))


def make_paste_case(idx: int, frame_hw=(1080, 1920), crop_res: int = 256):
    """One synthetic vehicle for the paste-back step (trajectory_inference.py:236-250): a bounding box (every fourth one
    hangs over a frame border so the square crop needs padding), the full-frame vehicle mask `dst_sketch_mask`
    (an ellipse inside the box, bool), and a `crop_res`^2 uint8 completed view.  Returns (bbox, mask, net_image);
    the crop geometry (`crop_info`) is derived from the bbox by the caller (utils/crop_utils.py:4-52)."""
    H, W = frame_hw
    rng = np.random.default_rng(70_000 + idx)
    bw = int(rng.integers(40, min(W, 480)))
    bh = int(rng.integers(30, min(H, 360)))
    if idx % 4 == 3:                                   # hug a border: the 1.1x square crop sticks out of the frame
        edge = (idx // 4) % 4
        x0 = 0 if edge == 0 else (W - 1 - bw if edge == 1 else int(rng.integers(0, W - bw)))
        y0 = 0 if edge == 2 else (H - 1 - bh if edge == 3 else int(rng.integers(0, H - bh)))
    else:
        x0 = int(rng.integers(0, W - bw))
        y0 = int(rng.integers(0, H - bh))
    x1, y1 = x0 + bw, y0 + bh
    yy, xx = np.mgrid[0:H, 0:W]
    cx, cy = (x0 + x1) / 2.0, (y0 + y1) / 2.0
    mask = ((xx - cx) / (bw / 2.0)) ** 2 + ((yy - cy) / (bh / 2.0)) ** 2 <= 1.0
    ys, xs = np.nonzero(mask)
    bbox = [int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())]          # as warp_learn/models.py:330-332
    net_image = rng.integers(0, 256, (crop_res, crop_res, 3), dtype=np.uint8)
    return bbox, mask, net_image


def make_pack_case(idx: int, frame_hw=(1080, 1920)):
    """One synthetic vehicle for VUNet input packing (trajectory_inference.py:205-227): the reference's
    `src_sketch_mask` (True = background), and source / destination normal sketches (uint8 RGB, black outside the
    vehicle, with a few black holes inside so the background-fill rule fires inside the box too)."""
    H, W = frame_hw
    bbox, veh, _ = make_paste_case(idx, frame_hw)
    rng = np.random.default_rng(80_000 + idx)
    yy, xx = np.mgrid[0:H, 0:W]

    def sketch(seed_shift):
        r = np.random.default_rng(81_000 + idx * 7 + seed_shift)
        n = np.stack([(xx * int(r.integers(1, 5)) + yy * int(r.integers(1, 5)) + int(r.integers(0, 255))) % 256 for _ in range(3)], -1).astype(np.uint8)
        n[~veh] = 0
        hx, hy = int(r.integers(bbox[0], bbox[2] + 1)), int(r.integers(bbox[1], bbox[3] + 1))
        n[max(hy - 6, 0):hy + 6, max(hx - 9, 0):hx + 9] = 0
        return n
    return np.logical_not(veh), sketch(0), sketch(1)


def make_trajectory_case(idx: int, steps: int = 20, h: int = 720, w: int = 1280):
    """One vehicle of the trajectory loop (trajectory_inference.py:255-367): CAD keypoints, a camera (K, R, t with
    X_cam = R X + t) that sees the vehicle, and `steps + 1` ground positions in metres (`meter_coords`, the output of
    trajectories_to_meters) along a gently curving path; every third vehicle turns sharply enough to trip the
    +-20 degree gates.  R is a plain look-at rotation; callers that need OpenCV's Rodrigues round trip apply it themselves."""
    rng = np.random.default_rng(9100 + idx)
    kp3d = cad_keypoints(idx % 10)
    K = np.array([[1100.0 + 20 * (idx % 5), 0, w / 2.0], [0, 1100.0 + 20 * (idx % 5), h / 2.0], [0, 0, 1.0]], np.float64)
    E = look_at_extrinsic(rng.uniform(0, 360), rng.uniform(10, 35), rng.uniform(14, 22))
    heading = rng.uniform(-np.pi, np.pi)
    turn = rng.uniform(-0.02, 0.02) if idx % 3 else rng.uniform(0.08, 0.16) * rng.choice([-1.0, 1.0])
    pos = [np.array([rng.uniform(-30, 30), rng.uniform(-30, 30)])]
    for n in range(steps):
        heading += turn + rng.normal(0, 0.01)
        if idx % 3 == 1 and n == steps // 2:
            heading += 0.6                                    # one abrupt change of direction: the "instant theta" gate
        pos.append(pos[-1] + rng.uniform(0.3, 0.7) * np.array([np.cos(heading), np.sin(heading)]))
    return dict(kp3d=kp3d, K=K, R=E[:3, :3].copy(), t=E[:3, 3].copy(), meter_coords=np.stack(pos), h=h, w=w)


def make_icn_pack_case(idx: int, frame_hw=(360, 640), res: int = 256):
    """One synthetic vehicle for `get_icn_inputs` (warp_learn/models.py:323-366): five full-frame warped planes (uint8 BGR,
    textured inside a per-plane polygon-ish region, zero elsewhere), the destination normal sketch (uint8 RGB) with its
    vehicle mask (True = vehicle) and a res x res central crop."""
    H, W = frame_hw
    bbox, veh, _ = make_paste_case(idx, frame_hw)
    rng = np.random.default_rng(83_000 + idx)
    yy, xx = np.mgrid[0:H, 0:W]
    normal = np.stack([(xx * int(rng.integers(1, 5)) + yy * int(rng.integers(1, 5)) + int(rng.integers(0, 255))) % 256 for _ in range(3)], -1).astype(np.uint8)
    normal[~veh] = 0
    planes = np.zeros((5, H, W, 3), np.uint8)
    for p in range(5):
        tex = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        band = veh & (((xx + yy * (p + 1)) // 23) % 5 == p)
        planes[p][band] = tex[band]
    central = rng.integers(0, 256, (res, res, 3), dtype=np.uint8)
    return planes, normal, veh, central


def make_car_mesh(cad_id: int = 0, n_lon: int = 64, n_lat: int = 24):
    """A closed synthetic "CAD" mesh for the normal-sketch renderer (no Pascal3D PLYs offline): a superellipsoid body the
    size of the box-car the keypoints of `cad_keypoints` sit on (x = left/right, y = back/front, z = up), with a raised
    cabin; per-id jitter like the keypoints.  Returns (vertices (Nv,3) f64, triangles (Nt,3) i32), outward oriented."""
    rng = np.random.default_rng(10_000 + cad_id)
    sx, sy, sz = 0.9 * (1 + rng.uniform(-0.05, 0.05)), 2.5 * (1 + rng.uniform(-0.05, 0.05)), 0.75 * (1 + rng.uniform(-0.05, 0.05))
    lat = np.linspace(-np.pi / 2, np.pi / 2, n_lat + 1)[1:-1]
    lon = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)

    def spow(t, e):
        return np.sign(t) * np.abs(t) ** e
    verts = [np.array([0.0, 0.0, -sz])]
    for la in lat:
        for lo in lon:
            x = sx * spow(np.cos(la), 0.5) * spow(np.cos(lo), 0.5)
            y = sy * spow(np.cos(la), 0.5) * spow(np.sin(lo), 0.5)
            z = sz * spow(np.sin(la), 0.6)
            if z > 0.3 * sz and abs(y) < 0.45 * sy:          # cabin
                z = z + 0.55 * sz * np.cos(y / (0.45 * sy) * np.pi / 2) ** 0.5
            verts.append(np.array([x, y, z]))
    verts.append(np.array([0.0, 0.0, sz + 0.55 * sz]))
    V = np.stack(verts)
    V[:, 2] += sz + 0.05                                      # wheels-on-ground: z >= 0
    tris = []
    R = len(lat)
    for k in range(n_lon):
        k1 = (k + 1) % n_lon
        tris.append((0, 1 + k1, 1 + k))
        for r in range(R - 1):
            a, b = 1 + r * n_lon + k, 1 + r * n_lon + k1
            c, d = a + n_lon, b + n_lon
            tris.append((a, b, d))
            tris.append((a, d, c))
        top = 1 + (R - 1) * n_lon
        tris.append((len(V) - 1, top + k, top + k1))
    return V, np.asarray(tris, np.int32)
