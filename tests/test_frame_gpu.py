"""Paste-back row: CUDA resize / paste-back against the numpy oracle and the cv2 goldens."""
import hashlib
import json
import os

import numpy as np
import pytest

from future_urban_scene_generation_b200 import synth

from oracle import frame_oracle as FO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "frame_golden.json")))


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_resize_matches_cv2_goldens(cuda):
    from future_urban_scene_generation_b200.frame_ops import resize_batch
    imgs, ds = [], []
    for c in GOLD["resize"]:
        sh, sw = c["src_hw"]
        imgs.append(np.random.default_rng(c["seed"]).integers(0, 256, (sh, sw, 3), dtype=np.uint8))
        ds.append((c["dst_hw"][1], c["dst_hw"][0]))
    outs = resize_batch(imgs, ds)                       # one ragged batch
    for c, o in zip(GOLD["resize"], outs):
        assert list(o.shape[:2]) == c["dst_hw"]
        assert sha(o.cpu().numpy()) == c["sha1"], c


def test_resize_matches_oracle_random_sizes(cuda):
    from future_urban_scene_generation_b200.frame_ops import resize_batch
    from oracle import frame_oracle as FO
    rng = np.random.default_rng(5)
    imgs, ds = [], []
    for t in range(60):
        sh, sw = (int(v) for v in rng.integers(1, 300, 2))
        dh, dw = (int(v) for v in rng.integers(1, 300, 2))
        if t % 6 == 0:
            sh, sw = 256, 256
        if t % 10 == 0:
            sh, sw = 2 * dh, 2 * dw                      # the exact-halving (box average) route
        imgs.append(rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8))
        ds.append((dw, dh))
    outs = resize_batch(imgs, ds)
    for im, d, o in zip(imgs, ds, outs):
        assert np.array_equal(o.cpu().numpy(), FO.resize_linear_u8(im, d)), (im.shape, d)


def _golden_sequence():
    Hf, Wf = GOLD["frame_hw"]
    frame = np.random.default_rng(GOLD["frame_seed"]).integers(0, 256, (Hf, Wf, 3), dtype=np.uint8)
    items = []
    for ci in GOLD["crop_info"]:
        bbox, mask, net = synth.make_paste_case(ci["idx"], (Hf, Wf))
        assert bbox == ci["bbox"]
        items.append((mask, net, {k: ci[k] for k in ("crop_xy_min", "pad_xy_before", "pad_xy_after", "crop_size_orig")}))
    return frame, items


def test_paste_back_sequence_matches_cv2_goldens(cuda):
    """Prefixes of the golden vehicle sequence pasted in ONE batched call each must land on the frame hash the
    sequential cv2 reference loop had after that many vehicles (overlapping masks: last vehicle wins)."""
    from future_urban_scene_generation_b200.frame_ops import paste_back_batch, paste_back
    frame, items = _golden_sequence()
    for n in (1, 5, 13, len(items)):
        out = paste_back_batch(frame[None].copy(), np.stack([it[1] for it in items[:n]]), [it[0] for it in items[:n]],
                               [it[2] for it in items[:n]], [0] * n)
        assert sha(out[0].cpu().numpy()) == GOLD["paste"][n - 1]["sha1_after"], n
    # one-vehicle drop-in, applied sequentially like the reference
    img = frame.copy()
    for k, (mask, net, info) in enumerate(items[:6]):
        paste_back(img, net, info, mask)
        assert sha(img) == GOLD["paste"][k]["sha1_after"]


def test_paste_back_multi_frame_and_rects_match_oracle(cuda):
    """Config-5-like: several frames, several vehicles per frame, masks given as bbox-sized sub-rectangles."""
    from future_urban_scene_generation_b200.frame_ops import paste_back_batch
    from oracle import frame_oracle as FO
    Hf, Wf, F, V = 270, 480, 3, 7
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (F, Hf, Wf, 3), dtype=np.uint8)
    ref = frames.copy()
    crops, masks, rects, infos, fidx = [], [], [], [], []
    for f in range(F):
        for v in range(V):
            bbox, mask, net = synth.make_paste_case(100 + f * V + v, (Hf, Wf))
            info = FO.square_crop_info((Hf, Wf), bbox)
            FO.paste_back(ref[f], net, info, mask)
            x0, y0, x1, y1 = bbox
            crops.append(net); infos.append(info); fidx.append(f)
            masks.append(mask[y0:y1 + 1, x0:x1 + 1]); rects.append((x0, y0, x1 - x0 + 1, y1 - y0 + 1))
    out = paste_back_batch(frames, np.stack(crops), masks, infos, fidx, mask_rects=rects)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_paste_back_rejects_what_the_reference_rejects(cuda):
    from future_urban_scene_generation_b200.frame_ops import paste_back_batch
    frame = np.zeros((1, 64, 64, 3), np.uint8)
    crop = np.zeros((1, 256, 256, 3), np.uint8)
    mask = np.ones((64, 64), bool)
    bad = {"crop_xy_min": (40, 0), "pad_xy_before": (0, 0), "pad_xy_after": (0, 0), "crop_size_orig": (30, 30)}   # 40 + 30 > 64
    with pytest.raises(ValueError):
        paste_back_batch(frame, crop, [mask], [bad], [0])


def test_pack_vunet_inputs_matches_reference_goldens(cuda):
    """x / y_tilde from the device == the reference lines (cv2 + the reference's helpers) recorded in the goldens, bit for
    bit; full-frame inputs (reference form) and bbox rectangles give the same tensors."""
    from future_urban_scene_generation_b200.frame_ops import pack_vunet_inputs_batch, pack_vunet_inputs
    Hf, Wf = GOLD["frame_hw"]
    frame = np.random.default_rng(GOLD["frame_seed"]).integers(0, 256, (Hf, Wf, 3), dtype=np.uint8)
    masks, ns, nd, rects, rm, rns, rnd = [], [], [], [], [], [], []
    for c in GOLD["pack"]:
        m, a, b = synth.make_pack_case(c["idx"], (Hf, Wf))
        masks.append(m); ns.append(a); nd.append(b)
        x0, y0, x1, y1 = c["bbox"]
        rects.append((x0, y0, x1 - x0 + 1, y1 - y0 + 1))
        rm.append(m[y0:y1 + 1, x0:x1 + 1]); rns.append(a[y0:y1 + 1, x0:x1 + 1]); rnd.append(b[y0:y1 + 1, x0:x1 + 1])
    n = len(masks)
    x, y, bbox = pack_vunet_inputs_batch(frame[None], [0] * n, masks, ns, nd)
    xr, yr, bboxr = pack_vunet_inputs_batch(frame[None], [0] * n, rm, rns, rnd, rects=rects)
    assert cuda.equal(x, xr) and cuda.equal(y, yr) and cuda.equal(bbox, bboxr)
    for i, c in enumerate(GOLD["pack"]):
        assert bbox[i].tolist() == c["bbox"]
        assert sha(x[i].cpu().numpy()) == c["sha1_x"], c["idx"]
        assert sha(y[i].cpu().numpy()) == c["sha1_y"], c["idx"]
    x1, y1 = pack_vunet_inputs(frame, masks[3], ns[3], nd[3])
    assert x1.shape == (1, 6, 256, 256) and y1.shape == (1, 3, 256, 256) and cuda.equal(x1[0], x[3]) and cuda.equal(y1[0], y[3])


def test_pack_vunet_inputs_feeds_the_model_like_host_tensors(cuda):
    """The packed device tensors are what the reference would hand to the VUNet: same completed crop as from the
    oracle's host-side packing."""
    from argparse import Namespace
    from future_urban_scene_generation_b200.frame_ops import pack_vunet_inputs_batch
    from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res
    from oracle import frame_oracle as FO
    torch = cuda
    Hf, Wf = 300, 500
    frame = np.random.default_rng(9).integers(0, 256, (Hf, Wf, 3), dtype=np.uint8)
    cases = [synth.make_pack_case(200 + i, (Hf, Wf)) for i in range(2)]
    x, y, _ = pack_vunet_inputs_batch(frame[None], [0, 0], [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases])
    ref = [FO.pack_vunet_inputs(frame, *c) for c in cases]
    assert np.array_equal(x.cpu().numpy(), np.stack([r[0] for r in ref])) and np.array_equal(y.cpu().numpy(), np.stack([r[1] for r in ref]))
    m = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).cuda().eval()
    torch.manual_seed(4)
    a = m(y, x)[0]
    torch.manual_seed(4)
    b = m(torch.from_numpy(np.stack([r[1] for r in ref])).cuda(), torch.from_numpy(np.stack([r[0] for r in ref])).cuda())[0]
    assert torch.equal(a, b)


def test_clip_chain_pack_vunet_paste_matches_oracle_chain(cuda):
    """The widened path end to end on the device -- pack inputs -> VUNet -> to_image -> paste back into the frames --
    against the same chain built from the CPU oracles (fp32 VUNet): the pasted pixels agree within the bf16 image
    tolerance (2 grey levels before the resize), everything outside the vehicle masks is untouched."""
    from argparse import Namespace
    from future_urban_scene_generation_b200.frame_ops import pack_vunet_inputs_batch, paste_back_batch
    from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res
    from future_urban_scene_generation_b200.warp_learn.planes_utils import to_image_batch, to_image
    from oracle import frame_oracle as FO, vunet_oracle as VO
    torch = cuda
    Hf, Wf, F, V = 240, 400, 2, 2
    frames = np.random.default_rng(21).integers(0, 256, (F, Hf, Wf, 3), dtype=np.uint8)
    sd = VO.make_state_dict(0)
    m = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    cases, fidx = [], []
    for f in range(F):
        for v in range(V):
            cases.append(synth.make_pack_case(300 + f * V + v, (Hf, Wf)))
            fidx.append(f)
    # device chain
    x, y, bbox = pack_vunet_inputs_batch(frames, fidx, [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases])
    torch.manual_seed(8)
    crops = to_image_batch(m(y, x)[0])
    infos = [FO.square_crop_info((Hf, Wf), bb) for bb in bbox.tolist()]          # crop_info of get_icn_inputs (same bbox rule)
    veh = [np.logical_not(c[0]) for c in cases]                                    # dst_sketch_mask after :176 = vehicle pixels
    out = paste_back_batch(frames.copy(), crops, veh, infos, fidx).cpu().numpy()
    # oracle chain
    ref = frames.copy()
    packed = [FO.pack_vunet_inputs(frames[f], *c) for f, c in zip(fidx, cases)]
    torch.manual_seed(8)
    with torch.no_grad():
        xt = VO.forward(sd, torch.from_numpy(np.stack([p[1] for p in packed])), torch.from_numpy(np.stack([p[0] for p in packed])))[0]
    for i, f in enumerate(fidx):
        FO.paste_back(ref[f], to_image(xt[i], from_LAB=False), infos[i], veh[i])
    union = np.zeros((F, Hf, Wf), bool)
    for i, f in enumerate(fidx):
        union[f] |= veh[i]
    assert np.array_equal(out[~union], frames[~union])                             # untouched outside the masks
    diff = np.abs(out.astype(int) - ref.astype(int))
    assert diff.max() <= 3 and (diff > 1).mean() < 0.01, (diff.max(), (diff > 1).mean())


def test_lab_conversion_matches_oracle_on_colour_sweep(cuda):
    """cv2's uint8 RGB -> Lab on the device (tables + exception list, exhaustively pinned to cv2 by scripts/make_lab_tables.py)
    vs the numpy oracle: every exception colour, a strided sweep of the colour cube and random colours, through
    fusg_pack_icn_inputs' central-crop channel (no resize involved)."""
    torch = cuda
    import os
    from future_urban_scene_generation_b200.frame_ops import get_icn_inputs_batch
    z = np.load(os.path.join(ROOT, "future_urban_scene_generation_b200", "data", "lab8.npz"))
    keys = z["exc_keys"].astype(np.int64)
    exc = np.stack([keys >> 16, (keys >> 8) & 255, keys & 255], -1).astype(np.uint8)
    sweep = np.stack(np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 3), np.arange(0, 256, 7), indexing="ij"), -1).reshape(-1, 3).astype(np.uint8)
    rnd = np.random.default_rng(5).integers(0, 256, (60000, 3), dtype=np.uint8)
    cols = np.concatenate([exc, sweep, rnd])
    n = 256 * 256
    B = (len(cols) + n - 1) // n
    cols = np.concatenate([cols, np.zeros((B * n - len(cols), 3), np.uint8)]).reshape(B, 256, 256, 3)
    planes = np.zeros((B, 5, 8, 8, 3), np.uint8)
    normal = np.zeros((B, 8, 8, 3), np.uint8)
    mask = np.zeros((B, 8, 8), bool)
    mask[:, 2:6, 2:6] = True
    out, _ = get_icn_inputs_batch(planes, normal, mask, cols)
    torch.cuda.synchronize()
    got = out[:, 3:6].cpu().numpy()
    for b in range(B):
        want = FO.lab_to_tensor(FO.rgb2lab_u8(cols[b]))
        assert np.array_equal(got[b].view(np.int32), want.view(np.int32)), b


def test_get_icn_inputs_matches_reference_goldens(cuda):
    """fusg_mask_bbox + fusg_pack_icn_inputs vs the reference's own get_icn_inputs (cv2.resize, cv2.cvtColor, PIL, torchvision)
    recorded by scripts/make_golden_icn_inputs.py, and vs the oracle bit for bit; the single-vehicle wrapper keeps the
    reference signature."""
    torch = cuda
    from future_urban_scene_generation_b200 import synth
    from future_urban_scene_generation_b200.frame_ops import get_icn_inputs_batch, get_icn_inputs
    for c in GOLD["icn_inputs"]:
        planes, normal, mask, central = synth.make_icn_pack_case(c["idx"], tuple(c["frame_hw"]))
        out, infos = get_icn_inputs_batch(planes[None], normal[None], mask[None], central[None])
        torch.cuda.synchronize()
        got = out[0].cpu().numpy()
        assert hashlib.sha1(np.ascontiguousarray(got).tobytes()).hexdigest() == c["sha1"], c["idx"]
        for k in ("crop_xy_min", "pad_xy_before", "pad_xy_after", "crop_size_orig"):
            assert list(infos[0][k]) == c[k], (c["idx"], k)
        want, winfo = FO.get_icn_inputs(planes, normal, mask, central)
        assert np.array_equal(got.view(np.int32), want.view(np.int32))
    # a batch of same-sized frames in one launch == item by item; planes may already live on the device
    cases = [synth.make_icn_pack_case(i, (360, 640)) for i in (0, 1, 3, 8)]
    out, infos = get_icn_inputs_batch(torch.from_numpy(np.stack([c[0] for c in cases])).cuda(), np.stack([c[1] for c in cases]),
                                      np.stack([c[2] for c in cases]), np.stack([c[3] for c in cases]))
    for i, c in enumerate(cases):
        one, info = get_icn_inputs(c[0], c[1], c[2], c[3], 256, 256)
        assert one.shape == (1, 21, 256, 256) and torch.equal(one[0], out[i]) and info == infos[i]


def test_warp_to_icn_chain_on_the_device(cuda):
    """The reference's real data flow (trajectory_inference.py:165-182): fused warp -> get_icn_inputs -> G_Resnet, without
    leaving the device, equals the oracle chain (bit-exact up to the generator's input, <= 1e-2 after it)."""
    torch = cuda
    from future_urban_scene_generation_b200 import synth
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    from future_urban_scene_generation_b200.warp_learn.models import G_Resnet
    from future_urban_scene_generation_b200.frame_ops import get_icn_inputs_batch
    from oracle import warp_oracle as WO, icn_oracle as IO
    B = 3
    batch = synth.make_warp_batch(40, B)
    res = warp_batch(batch["src"], batch["src_kp"], batch["dst_kp"], batch["K"], batch["E_src"], batch["E_dst"], batch["kp3d"])
    normals, masks, centrals = [], [], []
    for i in range(B):
        _, n, m, c = synth.make_icn_pack_case(20 + i, (256, 256))
        normals.append(n); masks.append(m); centrals.append(c)
    gen_in, infos = get_icn_inputs_batch(res.warped, np.stack(normals), np.stack(masks), np.stack(centrals))
    sd = IO.make_state_dict(0)
    g = G_Resnet(21)
    g.load_state_dict(sd, strict=True)
    g = g.cuda().eval()
    img = g(gen_in)
    torch.cuda.synchronize()
    import torch as T
    for i in range(B):
        w, _, _, _ = WO.warp_fused(batch["src"][i], batch["src_kp"][i], batch["dst_kp"][i], batch["K"][i], batch["E_src"][i], batch["E_dst"][i], batch["kp3d"][i])
        want_in, _ = FO.get_icn_inputs(w, normals[i], masks[i], centrals[i])
        assert np.array_equal(gen_in[i].cpu().numpy().view(np.int32), want_in.view(np.int32))
        with T.no_grad():
            want = IO.forward(sd, T.from_numpy(want_in)[None])
        assert (img[i].cpu() - want[0]).abs().max().item() <= 1e-2


def test_stacked_tensor_inputs_equal_per_item_lists(cuda):
    """The batched calls also take stacked full-frame tensors (already on the device) without per-item host work; results are
    those of the per-item list form."""
    torch = cuda
    from future_urban_scene_generation_b200.frame_ops import pack_vunet_inputs_batch, paste_back_batch
    Hf, Wf = 360, 640
    frame = np.random.default_rng(3).integers(0, 256, (1, Hf, Wf, 3), dtype=np.uint8)
    cases = [synth.make_pack_case(i, (Hf, Wf)) for i in range(5)]
    xl, yl, bl = pack_vunet_inputs_batch(frame, [0] * 5, [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases])
    xs, ys, bs = pack_vunet_inputs_batch(frame, [0] * 5, torch.from_numpy(np.stack([c[0] for c in cases])).cuda(),
                                         torch.from_numpy(np.stack([c[1] for c in cases])).cuda(), torch.from_numpy(np.stack([c[2] for c in cases])).cuda())
    assert torch.equal(xl, xs) and torch.equal(yl, ys) and torch.equal(bl, bs)
    items = [synth.make_paste_case(i, (Hf, Wf)) for i in range(6)]
    infos = [FO.square_crop_info((Hf, Wf), it[0]) for it in items]
    crops = np.stack([it[2] for it in items])
    fidx = [0, 1, 0, 1, 0, 1]
    base = np.random.default_rng(4).integers(0, 256, (2, Hf, Wf, 3), dtype=np.uint8)
    a = paste_back_batch(torch.from_numpy(base.copy()).cuda(), crops, [it[1] for it in items], infos, fidx)
    b = paste_back_batch(torch.from_numpy(base.copy()).cuda(), crops, torch.from_numpy(np.stack([it[1] for it in items])).cuda(), infos, fidx)
    assert torch.equal(a, b)


def test_to_image_from_lab_matches_oracle(cuda):
    """to_image(x, from_LAB=True) on the device (fusg_to_image_lab: float -> uint8, OpenCV's 8-bit Lab -> BGR) vs the numpy oracle,
    whose pipeline scripts/make_lab_tables.py verified against cv2 on all 2^24 (L, a, b) triples: a strided sweep of the Lab cube
    fed through exact float encodings, random network-like outputs incl. values outside [-1, 1], and the per-image mirror."""
    torch = cuda
    from future_urban_scene_generation_b200.warp_learn.planes_utils import to_image_batch, to_image
    sweep = np.stack(np.meshgrid(np.arange(0, 256, 3), np.arange(0, 256, 5), np.arange(0, 256, 5), indexing="ij"), -1).reshape(-1, 3)
    n = 256 * 256
    B = (len(sweep) + n - 1) // n
    lab = np.concatenate([sweep, np.zeros((B * n - len(sweep), 3), sweep.dtype)]).reshape(B, 256, 256, 3)
    x = np.transpose((lab.astype(np.float32) + 0.25) / 255 * 2 - 1, (0, 3, 1, 2)).copy()          # decodes back to `lab` after truncation
    got = to_image_batch(torch.from_numpy(x).cuda(), from_LAB=True).cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], FO.to_image_lab(x[b])), b
        assert np.array_equal(got[b], FO.lab2rgb_u8(lab[b].astype(np.uint8))[..., ::-1])
    rnd = np.random.default_rng(8).uniform(-1.3, 1.3, (3, 3, 128, 96)).astype(np.float32)
    got = to_image_batch(torch.from_numpy(rnd).cuda(), from_LAB=True).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], FO.to_image_lab(rnd[b]))
    assert np.array_equal(to_image(torch.from_numpy(rnd[0]).cuda(), from_LAB=True), got[0])
    assert np.array_equal(to_image(np.transpose(rnd[1], (1, 2, 0)), from_LAB=True), got[1])        # HWC ndarray form of the reference
