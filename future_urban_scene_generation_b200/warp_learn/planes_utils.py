"""Mirror of the reference's warp_learn/planes_utils.py: get_planes, warp_unwarp_planes,
planes_to_torch, to_image -- same names, arguments, return types and dtypes; the arithmetic
runs in libfusg.so on the GPU."""
from typing import List, Union

import numpy as np

from .. import _lib
from .online_visibility import pascal_texture_planes, _KP_NAMES


def _plane_vertices(kpoint_dict, pl_kp_names, h, w):
    # planes_utils.py:22-27 -- float normalised coords * (w,h), truncated by np.int32
    p2d = np.asarray([list(map(float, kpoint_dict[k])) for k in pl_kp_names])
    p2d[:, 0] *= w
    p2d[:, 1] *= h
    return np.int32(p2d)


def get_planes(image: np.ndarray, src_kpoint_dict, pascal_class: str, planes_visibility):
    """planes_utils.py:11-37 -> (planes (5,h,w,3) u8, [5 int32 arrays (6|4,2)], vis (5,) u8)."""
    torch = _lib.require_cuda()
    h, w = image.shape[:2]
    names = list(pascal_texture_planes[pascal_class].keys())
    kpoints_planes = [_plane_vertices(src_kpoint_dict, pascal_texture_planes[pascal_class][n], h, w) for n in names]
    # 12x2 table in _KP_NAMES order for the kernel (every keypoint belongs to some plane)
    kp12 = np.zeros((12, 2), np.int32)
    for n, arr in zip(names, kpoints_planes):
        for k, v in zip(pascal_texture_planes[pascal_class][n], arr):
            kp12[_KP_NAMES.index(k)] = v
    # keypoints outside the frame are served (cv2.fillPoly's clipped-edge rules are reproduced on the device);
    # the 16.16 fixed-point edge arithmetic is guaranteed up to |coordinate| <= 2^20 (csrc/warp_geom.cuh)
    if np.abs(kp12).max() > _lib.POLY_COORD_MAX:
        raise ValueError("get_planes: keypoint magnitude beyond 2^20 pixels")
    img = torch.as_tensor(np.ascontiguousarray(image, np.uint8)).cuda()
    kp = torch.as_tensor(kp12).cuda()
    planes = torch.empty((5, h, w, 3), dtype=torch.uint8, device="cuda")
    rc = _lib.lib().fusg_get_planes(_lib.ptr(img), _lib.ptr(kp), _lib.ptr(planes), 1, h, w, _lib.stream_ptr(torch))
    _lib.check(rc, "fusg_get_planes")
    visibilities = [planes_visibility[n] for n in names]
    return planes.cpu().numpy(), kpoints_planes, np.stack(visibilities).astype(np.uint8)


def warp_unwarp_planes(src_planes: np.ndarray, src_planes_kpoints: List[np.ndarray],
                       dst_planes_kpoints: List[np.ndarray], src_visibilities: np.ndarray,
                       dst_visibilities: np.ndarray, pascal_class: str, pascal_texture_planes):
    """planes_utils.py:40-82 on explicit plane images (the literal drop-in; the fused batch path is
    warp_learn.warp_batch).  Returns (planes_warped, planes_unwarped), uint8, shaped like src_planes."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    st = _lib.stream_ptr(torch)
    keys = list(pascal_texture_planes[pascal_class].keys())
    symmetry_set = [keys.index('left'), keys.index('right')]
    src_planes = np.ascontiguousarray(src_planes, np.uint8)
    n_pl, h, w = src_planes.shape[:3]
    planes_warped = np.zeros_like(src_planes)
    planes_unwarped = np.zeros_like(src_planes)
    # gating + symmetry remap (:57-68) on the host: five booleans
    todo = []
    for i in range(len(keys)):
        if not src_visibilities[i]:
            continue
        if i not in symmetry_set and not dst_visibilities[i]:
            continue
        if i in symmetry_set and 1 not in [dst_visibilities[j] for j in symmetry_set]:
            continue
        j = i
        if i in symmetry_set and not dst_visibilities[i]:
            j = symmetry_set[0] if i == symmetry_set[1] else symmetry_set[1]
        todo.append((i, j))
    if not todo:
        return planes_warped, planes_unwarped
    dev_src = torch.as_tensor(src_planes).cuda()
    for i, j in todo:
        s = np.ascontiguousarray(src_planes_kpoints[i], np.int32).reshape(-1, 2)
        d = np.ascontiguousarray(dst_planes_kpoints[j], np.int32).reshape(-1, 2)
        n = len(s)
        pts = torch.as_tensor(np.stack([np.stack([s, d]), np.stack([d, s])])).cuda()     # (2 dirs, 2, n, 2)
        srcs = pts[:, 0].contiguous()
        dsts = pts[:, 1].contiguous()
        Hm = torch.empty((2, 9), dtype=torch.float64, device="cuda")
        ok = torch.empty((2,), dtype=torch.uint8, device="cuda")
        _lib.check(L.fusg_find_homography(_lib.ptr(srcs), _lib.ptr(dsts), n, _lib.ptr(Hm), _lib.ptr(ok), 2, st),
                   "fusg_find_homography")
        if not bool(ok.all().item()):               # H12 is None or H21 is None (:74)
            continue
        warped = torch.empty((1, h, w, 3), dtype=torch.uint8, device="cuda")
        unwarped = torch.empty((1, h, w, 3), dtype=torch.uint8, device="cuda")
        _lib.check(L.fusg_warp_perspective(_lib.ptr(dev_src[i:i + 1].contiguous()), _lib.ptr(Hm[0:1].contiguous()),
                                           _lib.ptr(warped), 1, h, w, st), "fusg_warp_perspective")
        _lib.check(L.fusg_warp_perspective(_lib.ptr(warped), _lib.ptr(Hm[1:2].contiguous()), _lib.ptr(unwarped), 1, h, w, st),
                   "fusg_warp_perspective")
        planes_warped[j] = warped[0].cpu().numpy()
        planes_unwarped[i] = unwarped[0].cpu().numpy()
    return planes_warped, planes_unwarped


def planes_to_torch(planes, to_LAB: bool):
    """planes_utils.py:85-93: (5,h,w,3) u8 -> (5,3,h,w) f32 in [-1,1] (host tensor, like the reference)."""
    import torch
    if to_LAB:
        import cv2                                     # host colour conversion, ICN branch only
        planes = [cv2.cvtColor(p, cv2.COLOR_BGR2LAB) for p in planes]
    planes = np.stack([p for p in planes])
    planes = np.float32(planes) / 255.
    planes = np.transpose(planes, (0, 3, 1, 2))
    return (torch.from_numpy(planes) - 0.5) / 0.5


def to_image(x: Union[np.ndarray, "torch.Tensor"], from_LAB: bool):
    """planes_utils.py:96-118: [-1,1] CHW tensor (or HWC ndarray) -> uint8 HWC BGR (truncating cast)."""
    assert len(x.shape) == 3, f'Unsupported image shape {x.shape}'
    if from_LAB:
        # the ICN's Lab output: float -> uint8 and OpenCV's 8-bit Lab -> BGR on the device (fusg_to_image_lab)
        torch = _lib.require_cuda()
        t = x.detach() if hasattr(x, "detach") else torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(x, dtype=np.float32), (2, 0, 1))))
        return to_image_batch(t.float().cuda()[None], from_LAB=True)[0].cpu().numpy()
    try:
        x = x.to('cpu').detach().numpy()
        x = np.transpose(x, (1, 2, 0))
    except AttributeError:
        pass
    x = (x + 1.) / 2 * 255
    x = np.clip(x, 0, 255)
    x = x.astype(np.uint8)
    return x


_LAB_INV = {}


def to_image_batch(x, from_LAB=False):
    """to_image for a batch on the device: (B,3,H,W) fp32 CUDA tensor in [-1,1] -> (B,H,W,3) uint8 BGR CUDA tensor (same
    arithmetic as `to_image(..., from_LAB)`; libfusg.so: fusg_to_image, and fusg_to_image_lab with OpenCV's 8-bit Lab -> BGR
    for the ICN's Lab output)."""
    torch = _lib.require_cuda()
    assert x.is_cuda and x.dim() == 4 and x.shape[1] == 3
    x = x.float().contiguous()
    B, _, H, W = x.shape
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        if from_LAB:
            key = str(x.device)
            if key not in _LAB_INV:
                import os
                z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "lab8.npz"))
                _LAB_INV[key] = (torch.from_numpy(z["lab_to_yf"].astype(np.int32)).to(torch.uint16).to(x.device),
                                 torch.from_numpy(z["inv_gamma"]).to(x.device))
            yf, ig = _LAB_INV[key]
            _lib.check(_lib.lib().fusg_to_image_lab(_lib.ptr(x), _lib.ptr(out), _lib.ptr(yf), _lib.ptr(ig), B, H, W, _lib.stream_ptr(torch)),
                       "fusg_to_image_lab")
        else:
            _lib.check(_lib.lib().fusg_to_image(_lib.ptr(x), _lib.ptr(out), B, H, W, _lib.stream_ptr(torch)), "fusg_to_image")
    return out
