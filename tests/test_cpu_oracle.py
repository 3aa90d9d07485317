"""CPU suite, part 1: the oracle against OpenCV and against the committed reference goldens."""
import hashlib
import json
import os

import numpy as np
import pytest

from future_urban_scene_generation_b200 import synth
from oracle import warp_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
cv2 = pytest.importorskip("cv2")


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_fillpoly_matches_cv2_in_frame():
    rng = np.random.default_rng(0)
    for t in range(1500):
        H, W = (256, 256) if t % 2 else (96, 160)
        n = 4 if t % 3 else 6
        pts = np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], 1).astype(np.int32)
        if t % 7 == 0:
            pts[1] = pts[0]                    # duplicate vertex
        if t % 11 == 0:
            pts[:, 1] = pts[0, 1]              # all horizontal
        if t % 13 == 0:
            pts[:, 0] = pts[0, 0]              # all vertical
        ref = cv2.fillPoly(np.zeros((H, W), np.uint8), [pts], 1)
        got = O.fill_poly(np.zeros((H, W), np.uint8), pts, 1)
        assert np.array_equal(ref, got), pts.tolist()


def test_jacobi_eig_solve_invert_are_bit_exact():
    rng = np.random.default_rng(1)
    for t in range(100):
        n = 9 if t % 2 else 8
        A = rng.standard_normal((n, n))
        A = A @ A.T
        ok, w, v = cv2.eigen(A)
        W_, V_ = O.jacobi(A)
        assert np.array_equal(w.ravel(), W_) and np.array_equal(v, V_)
    for t in range(100):
        # the LM normal matrices: 12 x 9 Jacobians, badly scaled; every other one rank-deficient along one
        # direction (the scale gauge of the 9-parameter homography), which exercises the eigenvalue cut
        J = rng.standard_normal((12, 9)) * np.array([1e2, 1e2, 1, 1e2, 1e2, 1, 1e4, 1e4, 1e2])
        if t % 2:
            g = rng.standard_normal(9)
            J = J - np.outer(J @ g, g) / (g @ g)
        A = cv2.mulTransposed(J, True)
        b = rng.standard_normal(9)
        assert np.array_equal(cv2.solve(A, b.reshape(9, 1), flags=cv2.DECOMP_EIG)[1].ravel(), O.solve_eig(A, b))
        assert np.array_equal(cv2.invert(A, flags=cv2.DECOMP_EIG)[1], O.invert_eig(A))
        M = rng.standard_normal((3, 3))
        assert np.array_equal(cv2.invert(M)[1], O.invert3(M))


def _perspective_pair(rng, n):
    src = rng.integers(20, 236, (n, 2)).astype(np.int32)
    Ht = np.array([[1 + rng.uniform(-.2, .2), rng.uniform(-.2, .2), rng.uniform(-15, 15)],
                   [rng.uniform(-.2, .2), 1 + rng.uniform(-.2, .2), rng.uniform(-15, 15)],
                   [rng.uniform(-5e-4, 5e-4), rng.uniform(-5e-4, 5e-4), 1]])
    p = np.c_[src, np.ones(n)] @ Ht.T
    return src, np.int32(p[:, :2] / p[:, 2:] + rng.uniform(-1, 1, (n, 2)))


def test_findhomography_bit_exact():
    """4-point (DLT only) and 6-point (DLT + the 9-parameter LM refinement of opencv-python 4.13.0) results are
    bit-identical to cv2.findHomography; so is the None case."""
    rng = np.random.default_rng(2)
    for n in (4, 6):
        for t in range(400):
            src, dst = _perspective_pair(rng, n)
            Hc, _ = cv2.findHomography(src, dst)
            Ho = O.find_homography(src, dst)
            assert (Hc is None) == (Ho is None)
            if Hc is not None:
                assert np.array_equal(Hc, Ho), (n, t)
    # None <=> all src or all dst points share an x or a y
    for deg in range(4):
        src, dst = _perspective_pair(rng, 4)
        (src if deg < 2 else dst)[:, deg % 2] = 7
        assert cv2.findHomography(src, dst)[0] is None and O.find_homography(src, dst) is None
    # the side planes of synthetic cars, both directions
    for idx in range(3000, 3100):
        p = synth.make_pose_pair(idx)
        for pl in (0, 1):
            ids = O.plane_table(pl)
            for a_, b_ in ((p["src_kp"][ids], p["dst_kp"][ids]), (p["dst_kp"][ids], p["src_kp"][ids])):
                assert np.array_equal(cv2.findHomography(a_, b_)[0], O.find_homography(a_, b_)), (idx, pl)


def test_warpperspective_bit_exact_given_h():
    rng = np.random.default_rng(3)
    for t in range(8):
        H, W = (256, 256) if t % 2 else (90, 130)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        src, dst = _perspective_pair(rng, 4)
        Hm, _ = cv2.findHomography(np.minimum(src, [W - 1, H - 1]), np.minimum(dst, [W - 1, H - 1]))
        if Hm is None:
            continue
        ref = cv2.warpPerspective(img, Hm, dsize=(W, H))
        assert np.array_equal(ref, O.warp_perspective(img, Hm))


def test_oracle_reproduces_reference_goldens():
    """tests/golden/warp_golden.json was written by scripts/make_golden_warp.py from the imported
    reference (cv2 4.13.0): visibility, get_planes, plane_j and EVERY warped plane are identical."""
    gold = json.load(open(os.path.join(GOLD, "warp_golden.json")))
    H, W = gold["hw"]
    for case in gold["cases"][:60]:
        idx = case["idx"]
        p = synth.make_pose_pair(idx, H, W)
        img = synth.make_crop(idx, H, W)
        assert p["src_kp"].tolist() == case["src_kp"] and p["dst_kp"].tolist() == case["dst_kp"]
        warped, vis, pj, _ = O.warp_fused(img, p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
        assert vis[0].tolist() == case["vis_src"] and vis[1].tolist() == case["vis_dst"]
        assert pj.tolist() == case["plane_j"]
        planes = O.get_planes(img, p["src_kp"])
        assert [_sha(planes[j]) for j in range(5)] == case["planes_sha1"]
        assert [_sha(warped[j]) for j in range(5)] == case["warped_sha1"], idx


def test_fillpoly_matches_cv2_out_of_frame():
    """cv2.fillPoly's clipped-edge regime: vertices outside the canvas (outline drawn from the clipLine'd end points;
    interior edges take x -- always -- and y -- unless horizontal -- from the clipped segment)."""
    rng = np.random.default_rng(12)
    for t in range(4000):
        H, W = [(256, 256), (96, 160), (720, 1280), (64, 64)][t % 4]
        n = 4 if t % 2 else 6
        spread = [0.2, 0.6, 2.0, 8.0][(t // 4) % 4]
        c = np.array([rng.integers(-W // 4, W + W // 4), rng.integers(-H // 4, H + H // 4)])
        pts = (c + rng.normal(0, spread * max(H, W) / 4, (n, 2))).astype(np.int32)
        ref = cv2.fillPoly(np.zeros((H, W), np.uint8), [pts], 1)
        got = O.fill_poly(np.zeros((H, W), np.uint8), pts, 1)
        assert np.array_equal(ref, got), ((H, W), pts.tolist())


def test_oracle_reproduces_reference_out_of_frame_goldens():
    """Vehicles leaving the frame: the reference's outputs (hashes committed by scripts/make_golden_warp.py) at 256 x 256
    and at its own 1280 x 720 working resolution."""
    gold = json.load(open(os.path.join(GOLD, "warp_golden.json")))
    for key, (H, W), step in (("oob_sha1_16", (256, 256), 8), ("oob720_sha1_16", (720, 1280), 6)):
        for idx in range(0, len(gold[key]), step):
            p = synth.make_pose_pair(idx, H, W, out_of_frame=True)
            warped, vis, pj, _ = O.warp_fused(synth.make_crop(idx, H, W), p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
            got = hashlib.sha1(np.ascontiguousarray(vis[:2], np.uint8).tobytes() + np.ascontiguousarray(warped).tobytes()).hexdigest()[:16]
            assert got == gold[key][idx], (key, idx)


def test_oracle_reproduces_reference_bulk_goldens():
    """2000 further crops (seeds 1000..2999, 4268 written planes): sha1 over (visibility x2, warped planes) of the
    reference's outputs -- no tolerance, no exceptions."""
    gold = json.load(open(os.path.join(GOLD, "warp_golden.json")))
    H, W = gold["hw"]
    first = gold["bulk_first"]
    for k, want in enumerate(gold["bulk_sha1_16"][::4]):          # every 4th crop keeps the CPU suite short
        idx = first + 4 * k
        p = synth.make_pose_pair(idx, H, W)
        warped, vis, pj, _ = O.warp_fused(synth.make_crop(idx, H, W), p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
        got = hashlib.sha1(np.ascontiguousarray(vis[:2], np.uint8).tobytes() + np.ascontiguousarray(warped).tobytes()).hexdigest()[:16]
        assert got == want, idx


def test_fused_equals_stepwise_reference_order():
    """orc_warp_fused == compute_visibility x2 -> get_planes -> warp_unwarp_planes()[0]."""
    for idx in range(6):
        p = synth.make_pose_pair(idx, 128, 128)
        img = synth.make_crop(idx, 128, 128)
        warped, vis, pj, H12 = O.warp_fused(img, p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
        vs = O.compute_visibility(p["E_src"], p["K"], p["kp3d"], 128, 128)
        vd = O.compute_visibility(p["E_dst"], p["K"], p["kp3d"], 128, 128)
        sv = np.array([vs[n] for n in O.PLANE_NAMES[:5]], np.uint8)
        dv = np.array([vd[n] for n in O.PLANE_NAMES[:5]], np.uint8)
        planes = O.get_planes(img, p["src_kp"])
        w2, _, pj2, _ = O.warp_unwarp_planes(planes, p["src_kp"], p["dst_kp"], sv, dv)
        assert np.array_equal(warped, w2) and np.array_equal(pj, pj2)


def test_vunet_oracle_matches_golden_fingerprint():
    import torch
    from oracle import vunet_oracle as VO
    gold = json.load(open(os.path.join(GOLD, "vunet_golden.json")))
    sd = VO.make_state_dict(0)
    assert hashlib.sha1("\n".join(sd.keys()).encode()).hexdigest() == gold["key_sha1"] == "6a36f5dfc3dc8ab32fb79a78d759d8d28e09d940"
    assert sum(v.numel() for v in sd.values()) == gold["n_params"] == 45225158
    case = gold["cases"][0]
    x, y = synth.make_vunet_inputs(case["start"], case["B"])
    torch.manual_seed(case["noise_seed"])
    with torch.no_grad():
        o_x, o_mua, o_mus = VO.forward(sd, torch.from_numpy(y), torch.from_numpy(x))
    for name, t in (("x_tilde", o_x), ("mu_app0", o_mua[0]), ("mu_shape1", o_mus[1])):
        flat = t.flatten()
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        assert (flat[idx] - torch.tensor(case[name]["samples"])).abs().max().item() < 1e-4, name


def test_icn_oracle_matches_golden_fingerprint():
    """oracle/icn_oracle.py reproduces the fingerprints scripts/make_golden_icn.py recorded next to the imported reference
    (where it agreed with the reference G_Resnet to ~1e-6)."""
    import torch
    from oracle import icn_oracle as IO
    gold = json.load(open(os.path.join(GOLD, "icn_golden.json")))
    sd = IO.make_state_dict(0)
    assert hashlib.sha1("\n".join(sd.keys()).encode()).hexdigest() == gold["key_sha1"]
    assert IO.flops_per_crop(256) == gold["flops_per_crop_256"] == 130124087296          # 130.1 GFLOP/crop
    for case in gold["cases"][1:]:                                                        # the 64 / 128 px cases (seconds on CPU)
        assert max(case["oracle_vs_reference_maxabs"].values()) < 5e-5
        x = torch.from_numpy(synth.make_icn_inputs(case["start"], case["B"], case["res"]))
        with torch.no_grad():
            c = IO.encode(sd, x)
            o = IO.decode(sd, c)
        for name, t in (("out", o), ("content", c)):
            flat = t.flatten()
            idx = torch.linspace(0, flat.numel() - 1, 64).long()
            assert (flat[idx] - torch.tensor(case[name]["samples"])).abs().max().item() < 1e-4, name
        assert -1.0 <= case["out_range"][0] and case["out_range"][1] <= 1.0


def test_kinematics_oracle_matches_reference_goldens():
    """oracle/kinematics_oracle.py vs the reference lines run with numpy + cv2.projectPoints in the build container
    (scripts/make_golden_kinematics.py): SHA-256 of the moved keypoints, their projections and the int32 plane vertices."""
    from oracle import kinematics_oracle as KO
    from future_urban_scene_generation_b200 import kinematics as KM
    gold = json.load(open(os.path.join(GOLD, "kinematics_golden.json")))
    tripped = 0
    for g in gold["cases"]:
        case = synth.make_trajectory_case(g["idx"])
        R = np.array([float.fromhex(v) for v in g["R_cv_hex"]]).reshape(3, 3)
        theta, tr, rot = KO.trajectory_poses(case["meter_coords"])
        theta2, tr2, rot2 = KM.trajectory_poses(case["meter_coords"])          # the product's host-side mirror
        assert np.array_equal(theta, theta2) and np.array_equal(tr, tr2) and np.array_equal(rot, rot2)
        hm, h2, hv = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
        for s in range(g["steps"]):
            moved, kp2d, verts = KO.step(case["kp3d"], rot[s], tr[s], R, case["t"], case["K"], case["w"], case["h"])
            hm.update(moved.tobytes()); h2.update(kp2d.tobytes()); hv.update(verts.tobytes())
            if s == 0:
                assert verts.tolist() == g["first_verts"]
        assert hm.hexdigest() == g["sha256_moved"] and h2.hexdigest() == g["sha256_kp2d"] and hv.hexdigest() == g["sha256_verts"], g["idx"]
        assert [float(v).hex() for v in kp2d.flatten()] == g["last_kp2d_hex"]
        tripped += g["translation_gated_steps"]
    assert tripped >= 5                                    # the +-20 degree gates are exercised


def test_lab_and_icn_input_oracle_match_goldens():
    """oracle/frame_oracle.py: the uint8 Lab restatement against a few known OpenCV values and its own table invariants, and
    get_icn_inputs against the sha1 the reference function produced (scripts/make_golden_icn_inputs.py)."""
    from oracle import frame_oracle as FO
    z = np.load(os.path.join(os.path.dirname(GOLD), "..", "future_urban_scene_generation_b200", "data", "lab8.npz"))
    assert len(z["gamma_tab"]) == 256 and len(z["cbrt_tab"]) == 3072 and len(z["exc_keys"]) == len(z["exc_vals"]) == 1671
    assert np.all(np.diff(z["exc_keys"].astype(np.int64)) > 0)                       # sorted, unique: the device binary-searches it
    known = np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128]], np.uint8)
    assert FO.rgb2lab_u8(known).tolist() == [[0, 128, 128], [255, 128, 128], [136, 208, 195], [224, 42, 211], [82, 207, 20], [137, 128, 128]]
    known_lab = np.array([[0, 128, 128], [255, 128, 128], [136, 208, 195], [50, 0, 255], [200, 255, 0], [17, 90, 160]], np.uint8)
    assert FO.lab2rgb_u8(known_lab).tolist() == [[0, 0, 0], [255, 255, 255], [255, 2, 1], [0, 70, 0], [255, 47, 255], [0, 33, 0]]   # cv2 4.13 values
    assert len(z["lab_to_yf"]) == 512 and len(z["inv_gamma"]) == 4096 and z["inv_gamma"][0] == 0 and z["inv_gamma"][4095] == 255
    gold = json.load(open(os.path.join(GOLD, "frame_golden.json")))["icn_inputs"]
    for c in gold[:5]:
        planes, normal, mask, central = synth.make_icn_pack_case(c["idx"], tuple(c["frame_hw"]))
        got, info = FO.get_icn_inputs(planes, normal, mask, central)
        assert got.shape == (21, 256, 256) and got.dtype == np.float32
        assert hashlib.sha1(np.ascontiguousarray(got).tobytes()).hexdigest() == c["sha1"], c["idx"]
        assert list(info["crop_size_orig"]) == c["crop_size_orig"]


def test_space_depth_permutations_are_block_major():
    import torch
    from oracle import vunet_oracle as VO
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).view(2, 3, 4, 6)
    s = VO.space_to_depth(x)
    for dy in range(2):
        for dx in range(2):
            for c in range(3):
                assert torch.equal(s[:, (dy * 2 + dx) * 3 + c], x[:, c, dy::2, dx::2])
    assert torch.equal(VO.depth_to_space(s), x)


def test_frame_oracle_matches_cv2_goldens():
    """oracle/frame_oracle.py (cv2.resize INTER_LINEAR restatement + paste-back lines) against the vectors that
    scripts/make_golden_frame.py produced with cv2 and the reference's square_crop_from_bbox."""
    import hashlib
    import json
    from oracle import frame_oracle as FO
    from future_urban_scene_generation_b200 import synth
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "frame_golden.json")))
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
    for c in gold["resize"]:
        sh, sw = c["src_hw"]
        src = np.random.default_rng(c["seed"]).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert sha(FO.resize_linear_u8(src, (c["dst_hw"][1], c["dst_hw"][0]))) == c["sha1"], c
    Hf, Wf = gold["frame_hw"]
    img = np.random.default_rng(gold["frame_seed"]).integers(0, 256, (Hf, Wf, 3), dtype=np.uint8)
    for ci, p in zip(gold["crop_info"], gold["paste"]):
        bbox, mask, net = synth.make_paste_case(ci["idx"], (Hf, Wf))
        info = FO.square_crop_info((Hf, Wf), bbox)
        assert bbox == ci["bbox"]
        for k in ("crop_xy_min", "pad_xy_before", "pad_xy_after", "crop_size_orig"):
            assert list(info[k]) == ci[k], (ci["idx"], k)
        FO.paste_back(img, net, info, mask)
        assert sha(img) == p["sha1_after"], ci["idx"]
    frame = np.random.default_rng(gold["frame_seed"]).integers(0, 256, (Hf, Wf, 3), dtype=np.uint8)
    for c in gold["pack"][:8]:
        x, y, bbox = FO.pack_vunet_inputs(frame, *synth.make_pack_case(c["idx"], (Hf, Wf)))
        assert bbox == c["bbox"] and sha(x) == c["sha1_x"] and sha(y) == c["sha1_y"], c["idx"]


def test_render_oracle_properties():
    """The software restatement of the Open3D normal sketch (oracle/render_oracle.py): area-weighted vertex normals equal a
    direct per-triangle accumulation, the silhouette equals the union of the projected triangles, colours are the
    camera-independent normal colours, and a mesh rendered from the opposite side shows the opposite normals."""
    from oracle import render_oracle as RO
    V, T = synth.make_car_mesh(0, n_lon=24, n_lat=10)
    # vertex normals: unit length, outward on this convex-ish body
    n = RO.vertex_normals(V, T)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0)
    assert (np.sum(n * (V - V.mean(0)), 1) > 0).mean() > 0.95
    acc = np.zeros_like(V)
    tn = np.cross(V[T[:, 1]] - V[T[:, 0]], V[T[:, 2]] - V[T[:, 0]])
    for t in range(len(T)):
        acc[T[t]] += tn[t]
    assert np.allclose(n, acc / np.linalg.norm(acc, axis=1, keepdims=True), atol=1e-12)
    off, ids = RO.vertex_adjacency(T, len(V))
    assert off[-1] == 3 * len(T) and all(v in T[t] for v in (0, 5, len(V) - 1) for t in ids[off[v]:off[v + 1]])
    p = synth.make_pose_pair(3, 128, 128)
    img, mask, tri = RO.render_normals(V, T, p["E_src"], p["K"], 128, 128, return_ids=True)
    assert np.array_equal(mask, np.all(img == 0, axis=-1)) and (~mask).sum() > 300
    # silhouette == union of filled projected triangles (cv2 polygons, up to the 1-px outline cv2 adds)
    u, v, z = RO.project_vertices(V, p["E_src"], p["K"], 128, 128)
    sil = np.zeros((128, 128), np.uint8)
    for t in T:
        cv2.fillPoly(sil, [np.round(np.stack([u[t], v[t]], 1)).astype(np.int32)], 1)
    assert not ((~mask) & (sil == 0)).any()
    assert ((sil == 1) & mask).sum() < 0.2 * (sil == 1).sum()
    # visible surface faces the camera: decoded normal . (camera centre - point) > 0 for nearly all covered pixels
    cam = -p["E_src"][:3, :3].T @ p["E_src"][:3, 3]
    dec = img[~mask].astype(np.float64) / 255 * 2 - 1
    view = cam / np.linalg.norm(cam)
    assert (dec @ view > -0.15).mean() > 0.97
    # a rigid move of the mesh by the identity changes nothing; moving it out of view clears the frame
    img2, mask2 = RO.render_normals(V, T, p["E_src"], p["K"], 128, 128, rot=np.eye(3), tr=np.zeros(3))
    assert np.array_equal(img, img2)
    img3, mask3 = RO.render_normals(V, T, p["E_src"], p["K"], 128, 128, rot=np.eye(3), tr=np.array([500.0, 0, 0]))
    assert mask3.all()
