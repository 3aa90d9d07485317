"""GPU parity of the fused warp path (libfusg.so through the C ABI) against the CPU oracle.
Bar: bit-exact for visibility flags, plane indices, homography bits and warped pixels."""
import numpy as np
import pytest

from future_urban_scene_generation_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_batch(batch, h, w):
    from oracle import warp_oracle as O
    outs = [O.warp_fused(batch["src"][i], batch["src_kp"][i], batch["dst_kp"][i], batch["K"][i],
                         batch["E_src"][i], batch["E_dst"][i], batch["kp3d"][i]) for i in range(len(batch["src"]))]
    return [np.stack([o[k] for o in outs]) for k in range(4)]


@pytest.mark.parametrize("hw", [(256, 256), (128, 160), (100, 77), (300, 400), (720, 1280)])
def test_fused_matches_oracle(cuda, hw):
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    h, w = hw
    B = 40 if hw == (256, 256) else (12 if h <= 256 else 3)
    batch = synth.make_warp_batch(0, B, h, w)
    res = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
    cuda.cuda.synchronize()
    warped, vis, pj, H12 = _oracle_batch(batch, h, w)
    assert np.array_equal(res.vis.cpu().numpy(), vis)
    assert np.array_equal(res.plane_j.cpu().numpy(), pj)
    got_H = res.H12.cpu().numpy()
    assert np.array_equal(got_H.view(np.uint64), H12.view(np.uint64)), np.abs(got_H - H12).max()
    got = res.warped.cpu().numpy()
    diff = (got != warped)
    assert not diff.any(), f"{diff.sum()} differing bytes, planes {np.unique(np.nonzero(diff)[1])}"
    assert (pj >= 0).sum() >= B // 2     # the batch really exercises warped planes


def test_bulk_reference_goldens(cuda):
    """2000 crops (seeds 1000..2999, 4268 written planes): visibility of both poses and every warped plane of the
    CUDA path hash to what the imported reference (cv2 4.13.0) produced in the build container
    (scripts/make_golden_warp.py) -- bit-exact, no tolerance, including the LM-refined side planes."""
    import hashlib
    import json
    import os
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "warp_golden.json")))
    first, want = gold["bulk_first"], gold["bulk_sha1_16"]
    bad = []
    CH = 500
    for c0 in range(0, len(want), CH):
        batch = synth.make_warp_batch(first + c0, CH)
        res = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
        vis, warped = res.vis.cpu().numpy(), res.warped.cpu().numpy()
        for i in range(CH):
            got = hashlib.sha1(np.ascontiguousarray(vis[i, :2]).tobytes() + np.ascontiguousarray(warped[i]).tobytes()).hexdigest()[:16]
            if got != want[c0 + i]:
                bad.append(first + c0 + i)
    assert not bad, f"{len(bad)} of {len(want)} crops differ from the reference: {bad[:10]}"


@pytest.mark.parametrize("key,hw,chunk", [("oob_sha1_16", (256, 256), 500), ("oob720_sha1_16", (720, 1280), 12)])
def test_out_of_frame_reference_goldens(cuda, key, hw, chunk):
    """Vehicles leaving the frame (cv2.fillPoly's clipped-polygon regime, online_visibility.py:78-102 and
    planes_utils.py:25-31): every crop is warped, none refused, and the bytes equal the imported reference's."""
    import hashlib
    import json
    import os
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "warp_golden.json")))
    want = gold[key]
    h, w = hw
    bad, n_out = [], 0
    for c0 in range(0, len(want), chunk):
        n = min(chunk, len(want) - c0)
        batch = synth.make_warp_batch(c0, n, h, w, out_of_frame=True)
        n_out += int(((batch["src_kp"] < 0) | (batch["src_kp"] >= [w, h]) | (batch["dst_kp"] < 0) | (batch["dst_kp"] >= [w, h])).any(axis=(1, 2)).sum())
        res = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"], check=True)
        vis, warped = res.vis.cpu().numpy(), res.warped.cpu().numpy()
        assert not (res.plane_j == -2).any().item()
        for i in range(n):
            got = hashlib.sha1(np.ascontiguousarray(vis[i, :2]).tobytes() + np.ascontiguousarray(warped[i]).tobytes()).hexdigest()[:16]
            if got != want[c0 + i]:
                bad.append(c0 + i)
    assert n_out >= 0.9 * len(want)            # the set really is out of frame
    assert not bad, f"{len(bad)} of {len(want)} crops differ from the reference: {bad[:10]}"


def test_visibility_areas_match_oracle(cuda):
    from oracle import warp_oracle as O
    from future_urban_scene_generation_b200.warp_learn.online_visibility import compute_visibility_batch
    B = 64
    batch = synth.make_warp_batch(100, B, crops=False)
    vis, pts, areas = compute_visibility_batch(batch["E_src"], batch["K"], batch["kp3d"], 256, 256, return_aux=True)
    vis, pts, areas = vis.cpu().numpy(), pts.cpu().numpy(), areas.cpu().numpy()
    for i in range(B):
        d, opts = O.compute_visibility(batch["E_src"][i], batch["K"][i], batch["kp3d"][i], 256, 256, return_pts=True)
        assert np.array_equal(opts, pts[i])
        dist = O.plane_distances(batch["E_src"][i], batch["kp3d"][i])
        ovis, oareas = O.visibility_from_pts(opts, dist, 256, 256, return_areas=True)
        assert np.array_equal(oareas, areas[i]), (i, oareas, areas[i])
        assert np.array_equal(ovis, vis[i])


def test_compute_visibility_dropin(cuda):
    from oracle import warp_oracle as O
    from future_urban_scene_generation_b200.warp_learn.online_visibility import compute_visibility
    p = synth.make_pose_pair(7)
    kp3d = {k: p["kp3d"][i] for i, k in enumerate(synth.KP_NAMES)}
    got = compute_visibility(p["E_src"], p["K"], kp3d, 256, 256)
    assert list(got.keys()) == O.PLANE_NAMES
    assert got == O.compute_visibility(p["E_src"], p["K"], kp3d, 256, 256)
    # out-of-frame projections follow cv2's clipped polygons
    assert compute_visibility(p["E_src"], p["K"], kp3d, 64, 64) == O.compute_visibility(p["E_src"], p["K"], kp3d, 64, 64)


def test_get_planes_dropin(cuda):
    from oracle import warp_oracle as O
    from future_urban_scene_generation_b200.warp_learn.planes_utils import get_planes
    for idx, (h, w) in enumerate([(256, 256), (90, 130), (720, 1280)]):
        p = synth.make_pose_pair(idx, h, w)
        img = synth.make_crop(idx, h, w)
        kd = {k: p["kp2d_src"][i] for i, k in enumerate(synth.KP_NAMES)}
        vis = {n: bool(i % 2) for i, n in enumerate(O.PLANE_NAMES)}
        planes, kps, v = get_planes(img, kd, 'car', vis)
        assert planes.dtype == np.uint8 and planes.shape == (5, h, w, 3)
        assert np.array_equal(planes, O.get_planes(img, p["src_kp"]))
        for pl in range(5):
            assert kps[pl].dtype == np.int32
            assert np.array_equal(kps[pl], p["src_kp"][O.plane_table(pl)])
        assert v.dtype == np.uint8 and v.tolist() == [0, 1, 0, 1, 0]


def test_find_homography_bits(cuda):
    import ctypes
    from oracle import warp_oracle as O
    from future_urban_scene_generation_b200 import _lib
    torch = cuda
    rng = np.random.default_rng(3)
    for n in (4, 6):
        N = 300
        src = rng.integers(0, 256, (N, n, 2)).astype(np.int32)
        dst = np.clip(src + rng.integers(-25, 25, (N, n, 2)), 0, 255).astype(np.int32)
        src[0, :, 0] = 5                       # degenerate: all share an x -> None
        dst[1, :, 1] = 9
        ts, td = torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda()
        Hm = torch.empty((N, 9), dtype=torch.float64, device="cuda")
        ok = torch.empty((N,), dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib().fusg_find_homography(_lib.ptr(ts), _lib.ptr(td), n, _lib.ptr(Hm), _lib.ptr(ok), N,
                                                   _lib.stream_ptr(torch)), "fusg_find_homography")
        Hm, ok = Hm.cpu().numpy(), ok.cpu().numpy()
        for i in range(N):
            Ho = O.find_homography(src[i], dst[i])
            assert (Ho is not None) == bool(ok[i])
            if Ho is not None:
                same = np.array_equal(Ho.ravel().view(np.uint64), Hm[i].view(np.uint64)) or \
                    (np.isnan(Ho).any() and np.isnan(Hm[i]).any())
                assert same, (n, i, Ho.ravel(), Hm[i])
        assert ok[0] == 0 and ok[1] == 0


def test_warp_perspective_and_unwarp_dropin(cuda):
    from oracle import warp_oracle as O
    from future_urban_scene_generation_b200.warp_learn.planes_utils import get_planes, warp_unwarp_planes
    from future_urban_scene_generation_b200.warp_learn.online_visibility import pascal_texture_planes
    for idx in range(6):
        p = synth.make_pose_pair(idx)
        img = synth.make_crop(idx)
        planes = O.get_planes(img, p["src_kp"])
        skps = [p["src_kp"][O.plane_table(pl)] for pl in range(5)]
        dkps = [p["dst_kp"][O.plane_table(pl)] for pl in range(5)]
        sv = np.array([1, 1, 1, idx % 2, 1], np.uint8)
        dv = np.array([idx % 2, 1 - idx % 2, 1, 1, idx % 3 == 0], np.uint8)
        w_ref, u_ref, _, _ = O.warp_unwarp_planes(planes, p["src_kp"], p["dst_kp"], sv, dv)
        w_got, u_got = warp_unwarp_planes(planes, skps, dkps, sv, dv, 'car', pascal_texture_planes)
        assert w_got.dtype == np.uint8 and w_got.shape == planes.shape
        assert np.array_equal(w_got, w_ref)
        assert np.array_equal(u_got, u_ref)


def test_absurd_keypoints_are_refused_loudly(cuda):
    """Vertices beyond 2^20 px (a keypoint on the camera plane) are outside cv2's int32 rasteriser and the 16.16 edge
    arithmetic: the crop comes back refused (plane_j == -2, zero planes) and the host raises -- never silent."""
    from future_urban_scene_generation_b200.warp_learn import warp_batch, check_refused, RefusedCrops
    batch = synth.make_warp_batch(0, 4)
    batch["src_kp"][2, 3, 0] = (1 << 20) + 5
    res = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
    pj = res.plane_j.cpu().numpy()
    assert (pj[2] == -2).all() and (pj[[0, 1, 3]] != -2).all()
    assert not res.warped[2].any().item()
    with pytest.raises(RefusedCrops) as ei:
        check_refused(res)
    assert ei.value.indices == [2]
    with pytest.raises(RefusedCrops):
        warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"], check=True)


def test_get_planes_dropin_out_of_frame(cuda):
    from oracle import warp_oracle as O
    from future_urban_scene_generation_b200.warp_learn.planes_utils import get_planes
    for idx, (h, w) in enumerate([(256, 256), (90, 130), (720, 1280)] * 3):
        p = synth.make_pose_pair(idx, h, w, out_of_frame=True)
        img = synth.make_crop(idx, h, w)
        kd = {k: p["kp2d_src"][i] for i, k in enumerate(synth.KP_NAMES)}
        planes, kps, v = get_planes(img, kd, 'car', {n: True for n in O.PLANE_NAMES})
        assert np.array_equal(planes, O.get_planes(img, p["src_kp"]))


def test_large_batch_properties(cuda):
    """BASELINE config 3 shape at reduced count: untouched planes are exactly zero and every
    written plane is non-trivial; result is independent of batch composition."""
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    B = 512
    batch = synth.make_warp_batch(0, B)
    res = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
    pj = res.plane_j.cpu().numpy()
    written = np.zeros((B, 5), bool)
    for i in range(5):
        m = pj[:, i] >= 0
        written[np.nonzero(m)[0], pj[m, i]] = True
    nz = res.warped.reshape(B, 5, -1).any(dim=2).cpu().numpy()
    assert not (nz & ~written).any()
    assert (nz & written).sum() >= 0.95 * written.sum()
    sub = slice(100, 132)
    res2 = warp_batch(batch["src"][sub], batch["src_kp"][sub], batch["dst_kp"][sub], batch["K"][sub], batch["E_src"][sub],
                      batch["E_dst"][sub], batch["kp3d"][sub])
    assert cuda.equal(res2.warped, res.warped[sub])


def test_large_batch_solver_path_matches_small_batch(cuda):
    """Above 8192 (crop, plane) tasks the homographies are solved from the compacted task lists instead of one warp
    per (crop, plane); both launches must give the same bits (and the oracle's, by transitivity)."""
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    torch = cuda
    U, R = 96, 24                                   # 96 unique crops tiled to 2304 (> 8192 / 5)
    batch = synth.make_warp_batch(300, U)
    small = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
    big_in = {k: np.concatenate([v] * R, 0) for k, v in batch.items()}
    big = warp_batch(big_in["src"], big_in["src_kp"], big_in["dst_kp"], big_in["K"], big_in["E_src"], big_in["E_dst"], big_in["kp3d"])
    torch.cuda.synchronize()
    for r in (0, 7, R - 1):
        sl = slice(r * U, (r + 1) * U)
        assert torch.equal(big.plane_j[sl], small.plane_j)
        assert torch.equal(big.vis[sl], small.vis)
        assert torch.equal(big.H12[sl].view(torch.int64), small.H12.view(torch.int64))
        assert torch.equal(big.warped[sl], small.warped)


def test_large_batch_overwrites_stale_output(cuda):
    """The output contract is "fully written": every byte of a pre-dirtied output buffer must come out right on the
    large-batch (list) path, call after call -- the planes nobody warps into are zero-filled by the kernel itself."""
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    from future_urban_scene_generation_b200.warp_learn.batch import WarpResult
    torch = cuda
    U, R = 64, 30                                   # 1920 crops -> 9600 (crop, plane) tasks: the list path
    batch = synth.make_warp_batch(500, U)
    small = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
    big_in = {k: np.concatenate([v] * R, 0) for k, v in batch.items()}
    B = U * R
    out = WarpResult(warped=torch.full((B, 5, 256, 256, 3), 0xCD, dtype=torch.uint8, device="cuda"),
                     vis=torch.empty((B, 2, 7), dtype=torch.uint8, device="cuda"),
                     plane_j=torch.empty((B, 5), dtype=torch.int8, device="cuda"),
                     H12=torch.empty((B, 5, 3, 3), dtype=torch.float64, device="cuda"))
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in big_in.items()}
    for it in range(2):
        warp_batch(dev["src"], dev["src_kp"], dev["dst_kp"], dev["K"], dev["E_src"], dev["E_dst"], dev["kp3d"], out=out)
        torch.cuda.synchronize()
        for r in (0, 11, R - 1):
            sl = slice(r * U, (r + 1) * U)
            assert torch.equal(out.warped[sl], small.warped), (it, r)
            assert torch.equal(out.plane_j[sl], small.plane_j)
        out.warped.fill_(0x5A)
