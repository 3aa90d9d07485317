"""Mirror of the reference's warp_learn/online_visibility.py (public surface used by the hot
path: `pascal_texture_planes`, `compute_visibility`; reference file:line cited per symbol)."""
import numpy as np

from .. import _lib

# online_visibility.py:9-25 -- insertion order is part of the contract (plane index = position)
pascal_texture_planes = {
    'car': {
        'left': ['left_back_trunk', 'left_back_wheel', 'left_front_wheel',
                 'left_front_light', 'upper_left_windshield', 'upper_left_rearwindow'],
        'right': ['right_back_trunk', 'right_back_wheel', 'right_front_wheel',
                  'right_front_light', 'upper_right_windshield', 'upper_right_rearwindow'],
        'roof': ['upper_left_rearwindow', 'upper_left_windshield',
                 'upper_right_windshield', 'upper_right_rearwindow'],
        'front': ['left_front_light', 'right_front_light',
                  'upper_right_windshield', 'upper_left_windshield'],
        'back': ['left_back_trunk', 'right_back_trunk',
                 'upper_right_rearwindow', 'upper_left_rearwindow'],
    },
    'chair': {},
}

_KP_NAMES = ['left_back_trunk', 'left_back_wheel', 'left_front_light',
             'left_front_wheel', 'right_back_trunk', 'right_back_wheel',
             'right_front_light', 'right_front_wheel', 'upper_left_rearwindow',
             'upper_left_windshield', 'upper_right_rearwindow',
             'upper_right_windshield']                      # utils/keypoint_utils.py:9-13
_VIS_NAMES = ['left', 'right', 'roof', 'front', 'back', 'front_bt', 'back_bt']   # :110-114


def _extrinsic34(extrinsic):
    E = np.asarray(extrinsic, np.float64)
    assert E.shape == (3, 4) or E.shape == (4, 4)
    if E.shape == (4, 4):
        if not np.all(E[-1, :] == np.asarray([0, 0, 0, 1])):
            raise ValueError('Format for extrinsic not valid')          # online_visibility.py:46-49
        E = E[:3, :]
    return np.ascontiguousarray(E)


def compute_visibility_batch(extrinsics, intrinsics, kpoints_3d, h, w, return_aux=False):
    """B poses at once: extrinsics (B,3,4) f64, intrinsics (B,3,3), kpoints_3d (B,12,3) in
    _KP_NAMES order -> (B,7) uint8 on the device (0xff rows: a projection beyond 2^20 pixels, refused)."""
    torch = _lib.require_cuda()
    E = torch.as_tensor(np.ascontiguousarray(extrinsics, np.float64)).cuda()
    K = torch.as_tensor(np.ascontiguousarray(intrinsics, np.float64)).cuda()
    X = torch.as_tensor(np.ascontiguousarray(kpoints_3d, np.float64)).cuda()
    B = E.shape[0]
    vis = torch.empty((B, 7), dtype=torch.uint8, device="cuda")
    pts = torch.empty((B, 12, 2), dtype=torch.int32, device="cuda")
    areas = torch.empty((B, 7, 2), dtype=torch.int32, device="cuda")
    rc = _lib.lib().fusg_visibility(_lib.ptr(K), _lib.ptr(E), _lib.ptr(X), _lib.ptr(vis), _lib.ptr(pts),
                                    _lib.ptr(areas), B, int(h), int(w), _lib.stream_ptr(torch))
    _lib.check(rc, "fusg_visibility")
    return (vis, pts, areas) if return_aux else vis


def compute_visibility(extrinsic, intrinsic, kpoints_3d, h, w):
    """online_visibility.py:105-150.  extrinsic (3,4)|(4,4), intrinsic (3,3), kpoints_3d: dict
    name -> (3,) -> dict of 7 bools keyed left,right,roof,front,back,front_bt,back_bt."""
    E = _extrinsic34(extrinsic)[None]
    K = np.asarray(intrinsic, np.float64)
    assert K.shape == (3, 3)
    X = np.stack([np.asarray(kpoints_3d[k], np.float64).reshape(3) for k in _KP_NAMES])[None]
    vis, pts, _ = compute_visibility_batch(E, K[None], X, h, w, return_aux=True)
    vis = vis.cpu().numpy()[0]
    if vis[0] == 0xff:
        raise ValueError("compute_visibility: a keypoint projects beyond 2^20 pixels "
                         f"({pts.cpu().numpy()[0].tolist()}); is it on the camera plane?")
    return {n: bool(v) for n, v in zip(_VIS_NAMES, vis)}
