"""GPU parity of the per-step keypoint kinematics row through the C ABI: bit-exact against
oracle/kinematics_oracle.py (itself bit-identical to the reference lines executed with numpy + cv2.projectPoints,
tests/golden/kinematics_golden.json), and chained into the fused warp with the destination pose's own keypoints."""
import hashlib
import json
import os

import numpy as np
import pytest

from future_urban_scene_generation_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_step_keypoints_bit_exact_vs_oracle_and_goldens(cuda):
    torch = cuda
    from future_urban_scene_generation_b200 import kinematics as KM
    from oracle import kinematics_oracle as KO
    gold = json.load(open(os.path.join(GOLD, "kinematics_golden.json")))
    cases = [synth.make_trajectory_case(g["idx"]) for g in gold["cases"]]
    Rs = [np.array([float.fromhex(v) for v in g["R_cv_hex"]]).reshape(3, 3) for g in gold["cases"]]
    veh, rots, trs = [], [], []
    for v, c in enumerate(cases):
        _, tr, rot = KM.trajectory_poses(c["meter_coords"])
        veh += [v] * len(tr)
        rots.append(rot)
        trs.append(tr)
    h, w = cases[0]["h"], cases[0]["w"]
    moved, kp2d, verts = KM.step_keypoints_batch(np.stack([c["kp3d"] for c in cases]), np.array(veh), np.concatenate(rots), np.concatenate(trs),
                                                 np.stack(Rs), np.stack([c["t"] for c in cases]), np.stack([c["K"] for c in cases]), h, w)
    torch.cuda.synchronize()
    moved, kp2d, verts = moved.cpu().numpy(), kp2d.cpu().numpy(), verts.cpu().numpy()
    n = 0
    for v, (c, g) in enumerate(zip(cases, gold["cases"])):
        S = g["steps"]
        hm, h2, hv = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
        for s in range(S):
            hm.update(moved[n + s].tobytes()); h2.update(kp2d[n + s].tobytes()); hv.update(verts[n + s].tobytes())
        assert hm.hexdigest() == g["sha256_moved"] and h2.hexdigest() == g["sha256_kp2d"] and hv.hexdigest() == g["sha256_verts"], v
        for s in (0, S // 2, S - 1):                       # and against the oracle directly (bit patterns)
            o_m, o_2, o_v = KO.step(c["kp3d"], rots[v][s], trs[v][s], Rs[v], c["t"], c["K"], h, w)
            assert np.array_equal(moved[n + s].view(np.int64), o_m.view(np.int64))
            assert np.array_equal(kp2d[n + s].view(np.int64), o_2.view(np.int64)) and np.array_equal(verts[n + s], o_v)
        n += S
    assert n == len(veh) == 180


def test_trajectory_items_feed_the_fused_warp(cuda):
    """(vehicle, step) items -> fusg_step_keypoints -> fusg_warp_fused_traj equals the reference order run on the oracle:
    compute_visibility(E, kp3d) / compute_visibility(E, moved kp3d) -> get_planes -> warp_unwarp_planes()[0]."""
    torch = cuda
    from future_urban_scene_generation_b200 import kinematics as KM
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    from oracle import kinematics_oracle as KO, warp_oracle as O
    H = W = 256
    items = []
    for idx in range(40):
        p = synth.make_pose_pair(idx, H, W)
        rng = np.random.default_rng(4000 + idx)
        heading = rng.uniform(-np.pi, np.pi)
        mc = [np.zeros(2)]
        for _ in range(4):
            heading += rng.uniform(-0.05, 0.05)
            mc.append(mc[-1] + rng.uniform(0.1, 0.25) * np.array([np.cos(heading), np.sin(heading)]))
        _, tr, rot = KM.trajectory_poses(np.stack(mc))
        items.append((idx, p, tr, rot))
    V = len(items)
    veh = np.repeat(np.arange(V), 4)
    E = np.stack([it[1]["E_src"] for it in items])
    moved, kp2d, verts = KM.step_keypoints_batch(np.stack([it[1]["kp3d"] for it in items]), veh, np.concatenate([it[3] for it in items]),
                                                 np.concatenate([it[2] for it in items]), E[:, :3, :3], E[:, :3, 3],
                                                 np.stack([it[1]["K"] for it in items]), H, W)
    torch.cuda.synchronize()
    verts_h, moved_h = verts.cpu().numpy(), moved.cpu().numpy()
    inside = [(n, v) for n, v in enumerate(veh) if (verts_h[n] >= 0).all() and (verts_h[n, :, 0] < W).all() and (verts_h[n, :, 1] < H).all()]
    assert len(inside) >= 40
    sel = inside[:48]
    idxs = np.array([n for n, _ in sel])
    vs = np.array([v for _, v in sel])
    src = np.stack([synth.make_crop(items[v][0], H, W) for v in vs])
    res = warp_batch(src, np.stack([items[v][1]["src_kp"] for v in vs]), verts[torch.as_tensor(idxs).cuda()],
                     np.stack([items[v][1]["K"] for v in vs]), E[vs][:, :3], E[vs][:, :3], np.stack([items[v][1]["kp3d"] for v in vs]),
                     kp3d_dst=moved[torch.as_tensor(idxs).cuda()])
    torch.cuda.synchronize()
    warped, pj_d, vis_d = res.warped.cpu().numpy(), res.plane_j.cpu().numpy(), res.vis.cpu().numpy()
    differing_vis = 0
    for k, (n, v) in enumerate(sel):
        p = items[v][1]
        o_m, o_2, o_v = KO.step(p["kp3d"], items[v][3][n % 4], items[v][2][n % 4], p["E_src"][:3, :3], p["E_src"][:3, 3], p["K"], H, W)
        assert np.array_equal(o_v, verts_h[n]) and np.array_equal(o_m, moved_h[n])
        sd = O.compute_visibility(p["E_src"], p["K"], p["kp3d"], H, W)
        dd = O.compute_visibility(p["E_src"], p["K"], o_m, H, W)
        sv = np.array([sd[nm] for nm in O.PLANE_NAMES], np.uint8)
        dv = np.array([dd[nm] for nm in O.PLANE_NAMES], np.uint8)
        assert np.array_equal(vis_d[k, 0], sv) and np.array_equal(vis_d[k, 1], dv)
        differing_vis += int(not np.array_equal(sv, dv))
        w2, _, pj2, _ = O.warp_unwarp_planes(O.get_planes(src[k], p["src_kp"]), p["src_kp"], o_v, sv[:5], dv[:5])
        assert np.array_equal(pj_d[k], pj2) and np.array_equal(warped[k], w2), k
