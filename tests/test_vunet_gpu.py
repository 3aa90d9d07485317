"""GPU parity of the VUNet convolution engine (libfusg.so through the C ABI).

 * single convolutions: tcgen05 kernel vs the in-library direct kernel vs torch fp32 conv2d,
   over every shape class of the network (stride, 1x1/3x3, concat, 32/64/128/512 channels,
   DepthToSpace / SpaceToDepth / block scatter, Sampler noise, fp32 NCHW outputs);
 * whole network vs the fp32 oracle with identical weights and CPU noise:
   fp32 verification build <= 1e-4, bf16 product path <= 1e-2 (north_star tolerances).
"""
import ctypes as C
import json
import os
from argparse import Namespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_BF16 = 1e-2      # max-abs on [-1,1] images, north_star
TOL_FP32 = 1e-4


def _cfg():
    return Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)


# --------------------------------------------------------------------------- single convs
def _run_conv(torch, x_list, w, bias, stride, impl, dtype, residual=None, noise=None, mode=0, blk=0, want=("raw",)):
    """x_list: NCHW fp32 cuda tensors (1 or 2); w [cout,cin,k,k]; returns dict name -> NCHW fp32 tensor."""
    from future_urban_scene_generation_b200 import _lib
    from future_urban_scene_generation_b200.vunet.engine import ConvDesc
    L = _lib.lib()
    st = _lib.stream_ptr(torch)
    cd = 0 if dtype == "bf16" else 1
    td = torch.bfloat16 if dtype == "bf16" else torch.float32
    B, _, H, W = x_list[0].shape
    cout, cin, k, _ = w.shape
    cout_pad = (cout + 15) // 16 * 16
    keep = []

    def nhwc(t, cpad=None):
        t = t.float().contiguous()
        b, c, h, ww = t.shape
        cp = cpad or c
        o = torch.empty((b, h, ww, cp), dtype=td, device="cuda")
        _lib.check(L.fusg_nchw_to_nhwc(_lib.ptr(t), _lib.ptr(o), b, c, h, ww, cp, 0, cd, st), "nchw_to_nhwc")
        keep.append(t)
        return o
    ins = [nhwc(x, 32 if x.shape[1] < 16 else None) for x in x_list]
    cin_pad = sum(t.shape[-1] for t in ins)
    wv = torch.zeros((cout, cin_pad, k, k), device="cuda")
    # place real channels of each source at its (padded) offset
    off_r = off_p = 0
    for x, t in zip(x_list, ins):
        c = x.shape[1]
        wv[:, off_p:off_p + c] = w[:, off_r:off_r + c]
        off_r += c
        off_p += t.shape[-1]
    g = wv.flatten(1).norm(dim=1).contiguous()
    wp = torch.empty((cout_pad, k * k, cin_pad), dtype=td, device="cuda")
    _lib.check(L.fusg_fold_weightnorm(_lib.ptr(wv.contiguous()), _lib.ptr(g), _lib.ptr(wp), cout, cin_pad, k, cout_pad, cin_pad, cd, st), "fold")
    bp = torch.zeros((cout_pad,), device="cuda")
    bp[:cout] = bias
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    d = ConvDesc()
    d.in0, d.c0, d.pitch0 = ins[0].data_ptr(), ins[0].shape[-1], ins[0].shape[-1]
    if len(ins) > 1:
        d.in1, d.c1, d.pitch1 = ins[1].data_ptr(), ins[1].shape[-1], ins[1].shape[-1]
    d.B, d.H, d.W, d.ksize, d.stride = B, H, W, k, stride
    d.weight, d.bias, d.cout, d.cout_pad = wp.data_ptr(), bp.data_ptr(), cout, cout_pad
    if residual is not None:
        r = nhwc(residual)
        d.residual = r.data_ptr()
    if noise is not None:
        nz = noise.permute(0, 2, 3, 1).contiguous()
        d.noise = nz.data_ptr()
    if mode == 1:
        oshape = (B, 2 * Ho, 2 * Wo, cout // 4)
    elif mode == 2:
        oshape = (B, Ho // 2, Wo // 2, 4 * cout)
    elif mode == 3:
        oshape = (B, 2 * Ho, 2 * Wo, cout)
    else:
        oshape = (B, Ho, Wo, cout)
    outs = {}
    for i, name in enumerate(want):
        src = 1 if name.startswith("z") else 0
        elu = 1 if name.endswith("elu") else 0
        f32 = name.endswith("f32")
        if f32:
            t = torch.zeros((oshape[0], oshape[3], oshape[1], oshape[2]), dtype=torch.float32, device="cuda")
        else:
            t = torch.zeros(oshape, dtype=td, device="cuda")
        d.outs[i].ptr, d.outs[i].source, d.outs[i].elu, d.outs[i].layout = t.data_ptr(), src, elu, 1 if f32 else 0
        d.outs[i].mode, d.outs[i].blk = mode, blk
        outs[name] = t
    d.dtype, d.impl = cd, impl
    _lib.check(L.fusg_conv2d(C.byref(d), st), "fusg_conv2d")
    torch.cuda.synchronize()
    return {n: (t if n.endswith("f32") else t.float().permute(0, 3, 1, 2).contiguous()) for n, t in outs.items()}


def _torch_ref(torch, x_list, w, bias, stride, residual, noise, mode, blk):
    import torch.nn.functional as F
    from oracle import vunet_oracle as VO
    x = torch.cat(x_list, 1)
    y = F.conv2d(x.double(), w.double(), bias.double(), stride=stride, padding=w.shape[-1] // 2)
    if residual is not None:
        y = y + residual.double()
    z = y + noise.double() if noise is not None else None

    def place(t):
        if mode == 1:
            return VO.depth_to_space(t)
        if mode == 2:
            return VO.space_to_depth(t)
        if mode == 3:
            full = torch.zeros(t.shape[0], 4 * t.shape[1], t.shape[2], t.shape[3], dtype=t.dtype, device=t.device)
            full[:, blk * t.shape[1]:(blk + 1) * t.shape[1]] = t
            return VO.depth_to_space(full)
        return t
    return place(y), (place(z) if z is not None else None)


CONV_CASES = [
    # name, B, cins, cout, k, stride, H, residual, noise, mode
    ("res128_s1", 2, (128,), 128, 3, 1, 16, True, False, 0),
    ("down128_s2", 2, (128,), 128, 3, 2, 32, False, False, 0),
    ("nin128_1x1", 3, (128,), 128, 1, 1, 8, False, False, 0),
    ("cat256_res", 2, (128, 128), 128, 3, 1, 16, True, False, 0),
    ("first_6ch", 1, (6,), 128, 1, 1, 32, False, False, 0),
    ("res32", 1, (32,), 32, 3, 1, 64, True, False, 0),
    ("cat64_to32", 1, (32, 32), 32, 3, 1, 64, True, False, 0),
    ("down32_64", 2, (32,), 64, 3, 2, 32, False, False, 0),
    ("res64", 2, (64,), 64, 3, 1, 32, True, False, 0),
    ("up_d2s_512", 2, (128,), 512, 3, 1, 8, False, False, 1),
    ("up_d2s_128", 1, (32,), 128, 3, 1, 32, False, False, 1),
    ("res_s2d", 5, (128,), 128, 3, 1, 4, True, False, 2),
    ("ar_res_1024", 5, (512, 512), 512, 3, 1, 2, True, False, 0),
    ("sampler_blk", 5, (512,), 128, 3, 1, 2, False, True, 3),
    ("ar_res_1024_b64", 64, (512, 512), 512, 3, 1, 2, True, False, 0),   # AR-block shapes at the bench batch size
    ("sampler_blk_b64", 64, (512,), 128, 3, 1, 4, False, True, 3),
    ("sampler_plain", 3, (128,), 128, 3, 1, 4, False, True, 0),
    ("final_3ch", 1, (32,), 3, 3, 1, 64, False, False, 0),
    ("big_tile_256", 1, (128,), 128, 3, 1, 256, True, False, 0),
    ("msub2_wide", 4, (128,), 128, 3, 1, 256, True, False, 0),        # enough tiles for 256-row CTA tiles
    ("msub2_down", 10, (128,), 128, 3, 2, 256, False, False, 0),     # 640 256-row tiles >= 4 x 148: stride 2 WITH msub = 2
    ("msub2_down_64", 40, (128,), 128, 3, 2, 128, False, False, 0),  # the 128^2 -> 64^2 down-sampler at bench-like tile counts
    ("pair_w256", 3, (128,), 128, 3, 1, 256, True, False, 0),        # cta_group::2 pairs, two tile columns
    ("pair_w128", 10, (128,), 128, 3, 1, 128, True, False, 0),       # cta_group::2 pairs on a one-tile-column frame (W = N = 128)
    ("msub2_cat32", 4, (32, 32), 32, 3, 1, 256, True, False, 0),
    ("halo_cat64", 4, (64, 64), 64, 3, 1, 256, True, False, 0),       # sliding-window A tiles, streamed weights
    ("halo_res64", 4, (64,), 64, 3, 1, 256, True, False, 0),          # sliding-window A tiles, resident weights
    ("halo_w128", 16, (64,), 64, 3, 1, 128, True, False, 0),          # one tile column (image row == sub-tile)
]


def test_smallcin_streaming_nin(cuda):
    """FUSG_IMPL_SMALLCIN (the first NiN of each encoder: 6 / 3 real channels stored 16 wide, 1x1): selected automatically,
    equal to the tensor-core kernel on the same descriptor and to torch on the bf16-rounded operands."""
    torch = cuda
    import ctypes as C
    from future_urban_scene_generation_b200 import _lib
    from future_urban_scene_generation_b200.vunet.engine import ConvDesc
    g = torch.Generator(device="cpu").manual_seed(12)
    for cin, cout, B, H in ((6, 128, 3, 64), (3, 32, 2, 128)):
        x = torch.randn((B, H, H, cin), generator=g)
        xin = torch.zeros((B, H, H, 16), dtype=torch.bfloat16, device="cuda")
        xin[..., :cin] = x.cuda().bfloat16()
        w = torch.zeros((cout, 1, 64), dtype=torch.bfloat16, device="cuda")
        w[:, 0, :cin] = (torch.randn((cout, cin), generator=g) / cin ** 0.5).cuda().bfloat16()
        bias = (torch.randn((cout,), generator=g) * 0.1).cuda()
        outs = {}
        for impl in (0, 1):
            raw = torch.empty((B, H, H, cout), dtype=torch.bfloat16, device="cuda")
            elu = torch.empty_like(raw)
            d = ConvDesc()
            d.in0, d.c0, d.pitch0, d.cphys0, d.cin_real = xin.data_ptr(), 64, 16, 16, cin
            d.B, d.H, d.W, d.ksize, d.stride = B, H, H, 1, 1
            d.weight, d.bias, d.cout, d.cout_pad = w.data_ptr(), bias.data_ptr(), cout, cout
            d.outs[0].ptr, d.outs[1].ptr, d.outs[1].elu = raw.data_ptr(), elu.data_ptr(), 1
            d.dtype, d.impl = 0, impl
            if impl == 0:
                assert _lib.lib().fusg_conv2d_select(C.byref(d)) == 3
            _lib.check(_lib.lib().fusg_conv2d(C.byref(d), _lib.stream_ptr(torch)), "fusg_conv2d")
            torch.cuda.synchronize()
            outs[impl] = (raw.float(), elu.float())
        ref = torch.einsum("bhwc,oc->bhwo", xin[..., :cin].float(), w[:, 0, :cin].float()) + bias
        for impl in (0, 1):
            assert (outs[impl][0] - ref).abs().max().item() < 1.0 / 128 * max(1.0, ref.abs().max().item())
            assert (outs[impl][1] - torch.nn.functional.elu(outs[impl][0])).abs().max().item() < 1.0 / 128
        assert (outs[0][0] - outs[1][0]).abs().max().item() < 1.0 / 64 * max(1.0, ref.abs().max().item())


# the kernel variant a case is named after must really be the one that ran (fusg_conv2d_last_plan)
EXPECTED_PLAN = {
    "msub2_wide": {"msub": 2, "halo": 1, "pair": 1},
    "msub2_down": {"msub": 2, "halo": 0, "pair": 0},
    "msub2_down_64": {"msub": 2, "halo": 0},
    "pair_w256": {"msub": 2, "pair": 1},
    "pair_w128": {"msub": 2, "pair": 1},
    "halo_cat64": {"halo": 1},
    "halo_res64": {"halo": 1, "w_resident": 1},
    "halo_w128": {"halo": 1},
    "ar_res_1024": {"msub": 1},
}


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_kernels(cuda, case):
    torch = cuda
    name, B, cins, cout, k, stride, H, use_res, use_noise, mode = case
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000)
    xs = [torch.randn((B, c, H, H), generator=g).cuda() for c in cins]
    cin = sum(cins)
    w = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.randn((cout,), generator=g).cuda() * 0.1
    Ho = (H - 1) // stride + 1
    res = torch.randn((B, cout, Ho, Ho), generator=g).cuda() if use_res else None
    noise = torch.randn((B, cout, Ho, Ho), generator=g).cuda() if use_noise else None
    blk = 2
    want = ("raw", "elu", "rawf32") + (("zraw", "zelu", "zf32") if use_noise else ())
    if cout == 3:
        want = ("rawf32",)
    y_ref, z_ref = _torch_ref(torch, xs, w, bias, stride, res, noise, mode, blk)
    # fp32 direct kernel against torch: the verification path
    o32 = _run_conv(torch, xs, w, bias, stride, 2, "fp32", res, noise, mode, blk, want)
    assert (o32["rawf32"].double() - y_ref).abs().max().item() < 2e-5
    if "raw" in o32:
        assert (o32["raw"].double() - y_ref).abs().max().item() < 2e-5
        assert (o32["elu"].double() - torch.nn.functional.elu(y_ref)).abs().max().item() < 2e-5
    if use_noise:
        assert (o32["zf32"].double() - z_ref).abs().max().item() < 2e-5
    # bf16 inputs: reference with the same rounded operands
    xs_b = [x.bfloat16().float() for x in xs]
    w_b = w.bfloat16().float()
    res_b = res.bfloat16().float() if res is not None else None
    y_b, z_b = _torch_ref(torch, xs_b, w_b, bias, stride, res_b, noise, mode, blk)
    scale = max(1.0, y_b.abs().max().item())
    for impl in (2, 1):
        ob = _run_conv(torch, xs, w, bias, stride, impl, "bf16", res, noise, mode, blk, want)
        if impl == 1 and name in EXPECTED_PLAN:
            from future_urban_scene_generation_b200 import _lib
            plan = _lib.conv_last_plan()
            for key, val in EXPECTED_PLAN[name].items():
                assert plan[key] == val, (name, plan)
        err = (ob["rawf32"].double() - y_b).abs().max().item()
        assert err < 2e-3 * scale, (name, impl, "f32 out", err)      # only fp32 accumulation order differs
        if "raw" in ob:
            err = (ob["raw"].double() - y_b).abs().max().item()
            assert err < 1.0 / 128 * scale, (name, impl, "bf16 out", err)
            err = (ob["elu"].double() - torch.nn.functional.elu(y_b)).abs().max().item()
            assert err < 1.0 / 64 * scale, (name, impl, "elu out", err)
        if use_noise:
            err = (ob["zf32"].double() - z_b).abs().max().item()
            assert err < 2e-3 * scale
            err = (ob["zraw"].double() - z_b).abs().max().item()
            assert err < 1.0 / 128 * max(scale, z_b.abs().max().item())


# --------------------------------------------------------------------------- whole network
@pytest.fixture(scope="module")
def nets(cuda):
    from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res
    from oracle import vunet_oracle as VO
    sd = VO.make_state_dict(0)
    m = Vunet_fix_res(_cfg())
    assert hasattr(m.load_state_dict(sd, strict=True), "missing_keys")
    m = m.to("cuda").eval()
    return m, sd, VO


def _inputs(torch, start, B):
    from future_urban_scene_generation_b200 import synth
    x, y = synth.make_vunet_inputs(start, B)
    return torch.from_numpy(x), torch.from_numpy(y)


def _maxabs(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


@pytest.mark.parametrize("dtype,impl,tol", [("fp32", "direct", TOL_FP32), ("bf16", "direct", TOL_BF16), ("bf16", "auto", TOL_BF16)])
def test_forward_matches_oracle(cuda, nets, dtype, impl, tol):
    torch = cuda
    m, sd, VO = nets
    m.set_compute(dtype, impl)
    x, y = _inputs(torch, 0, 2)
    torch.manual_seed(1)
    with torch.no_grad():
        r_x, r_mua, r_mus = VO.forward(sd, y, x)
    torch.manual_seed(1)
    g_x, g_mua, g_mus = m(y.cuda(), x.cuda())
    torch.cuda.synchronize()
    assert g_x.shape == (2, 3, 256, 256) and g_x.dtype == torch.float32 and g_x.is_cuda
    errs = {"x_tilde": _maxabs(g_x, r_x)}
    for i in range(2):
        assert g_mua[i].shape == r_mua[i].shape and g_mus[i].shape == r_mus[i].shape
        errs[f"mu_app{i}"] = _maxabs(g_mua[i], r_mua[i])
        errs[f"mu_shape{i}"] = _maxabs(g_mus[i], r_mus[i])
    print(dtype, impl, errs)
    assert max(errs.values()) <= tol, errs


def test_forward_matches_oracle_batch8(cuda, nets):
    """Batch 8 is the smallest batch at which the wide layers take the large-batch kernel variants of the bench
    (256-row CTA tiles, sliding-window A tiles, warp-staged epilogue); batch 2 above never reaches them."""
    torch = cuda
    m, sd, VO = nets
    m.set_compute("bf16", "auto")
    x, y = _inputs(torch, 40, 8)
    torch.manual_seed(3)
    with torch.no_grad():
        r_x, r_mua, r_mus = VO.forward(sd, y, x)
    torch.manual_seed(3)
    g_x, g_mua, g_mus = m(y.cuda(), x.cuda())
    torch.cuda.synchronize()
    errs = {"x_tilde": _maxabs(g_x, r_x)}
    for i in range(2):
        errs[f"mu_app{i}"] = _maxabs(g_mua[i], r_mua[i])
        errs[f"mu_shape{i}"] = _maxabs(g_mus[i], r_mus[i])
    print("batch8", errs)
    assert max(errs.values()) <= TOL_BF16, errs


def test_golden_fingerprint(cuda, nets):
    """The committed fingerprints were produced by the oracle after it was checked against the
    imported reference (scripts/make_golden_vunet.py); the CUDA path must land on them too."""
    torch = cuda
    m, sd, VO = nets
    m.set_compute("bf16", "auto")
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vunet_golden.json")))
    case = gold["cases"][0]
    x, y = _inputs(torch, case["start"], case["B"])
    torch.manual_seed(case["noise_seed"])
    g_x, g_mua, g_mus = m(y.cuda(), x.cuda())
    for name, t in (("x_tilde", g_x), ("mu_app0", g_mua[0]), ("mu_app1", g_mua[1]), ("mu_shape0", g_mus[0]), ("mu_shape1", g_mus[1])):
        flat = t.detach().float().cpu().flatten()
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        want = torch.tensor(case[name]["samples"])
        assert list(t.shape) == case[name]["shape"]
        assert (flat[idx] - want).abs().max().item() <= TOL_BF16, name


def test_subforward_api_traj_style(cuda, nets):
    """trajectory_inference.py:230-233: enc_up -> enc_down -> dec_up -> dec_down(..., mu_app)."""
    torch = cuda
    m, sd, VO = nets
    m.set_compute("bf16", "auto")
    x, y = _inputs(torch, 3, 1)
    torch.manual_seed(2)
    with torch.no_grad():
        oe, se = VO.forward_enc_up(sd, x)
        mu_r, z_r = VO.forward_enc_down(sd, oe, se)
        od, sdn = VO.forward_dec_up(sd, y)
        img_r, mu2_r, z2_r = VO.forward_dec_down(sd, od, sdn, mu_r)
    torch.manual_seed(2)
    oe_g, se_g = m.forward_enc_up(x.cuda())
    assert len(oe_g) == 2 and len(se_g) == 2 and oe_g[0].shape == (1, 128, 4, 4) and se_g[0].shape == (1, 128, 8, 8)
    mu_g, z_g = m.forward_enc_down(oe_g, se_g)
    od_g, sd_g = m.forward_dec_up(y.cuda())
    assert len(od_g) == 1 and len(sd_g) == 14 and sd_g[0].shape == (1, 32, 256, 256) and sd_g[-1].shape == (1, 128, 4, 4)
    for a, b in zip(sd_g, sdn + []):
        pass
    img_g, mu2_g, z2_g = m.forward_dec_down(od_g, sd_g, mu_g)
    assert sd_g == []                                   # popped empty like the reference
    errs = [_maxabs(img_g, img_r)] + [_maxabs(a, b) for a, b in zip(mu_g + z_g + mu2_g + z2_g, mu_r + z_r + mu2_r + z2_r)]
    errs += [_maxabs(a, b) for a, b in zip(oe_g + se_g + od_g, oe + se + od)]
    print("traj-style errs", errs)
    assert max(errs) <= TOL_BF16
    # foreign tensors (clones carry no engine tag) must give the same answer
    torch.manual_seed(2)
    oe_g, se_g = m.forward_enc_up(x.cuda())
    mu_f, z_f = m.forward_enc_down([t.clone() for t in oe_g], [t.clone() for t in se_g])
    od_g, sd_g = m.forward_dec_up(y.cuda())
    img_f, _, _ = m.forward_dec_down([t.clone() for t in od_g], [t.clone() for t in sd_g], [t.clone() for t in mu_f])
    assert _maxabs(img_f, img_r) <= TOL_BF16


def test_mean_shape_mode(cuda, nets):
    torch = cuda
    m, sd, VO = nets
    m.set_compute("bf16", "auto")
    x, y = _inputs(torch, 9, 1)
    torch.manual_seed(3)
    with torch.no_grad():
        r = VO.forward(sd, y, None, mean_mode="mean_shape")
    torch.manual_seed(3)
    g = m(y.cuda(), None, mean_mode="mean_shape")
    assert _maxabs(g, r) <= TOL_BF16


def test_state_dict_roundtrip_and_refold(cuda, nets):
    torch = cuda
    m, sd, VO = nets
    m.set_compute("bf16", "auto")
    back = m.state_dict()
    assert list(back.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(back[k].cpu(), sd[k])
    x, y = _inputs(torch, 0, 1)
    torch.manual_seed(1)
    a = m(y.cuda(), x.cuda())[0].clone()
    sd2 = {k: (v * 1.5 if k.endswith("shape_decoder_6.conv.conv.weight_g") else v) for k, v in sd.items()}
    m.load_state_dict(sd2, strict=True)
    torch.manual_seed(1)
    b = m(y.cuda(), x.cuda())[0].clone()
    assert (a - b).abs().max().item() > 1e-3          # new weights were re-folded
    m.load_state_dict(sd, strict=True)
    torch.manual_seed(1)
    c = m(y.cuda(), x.cuda())[0]
    assert torch.equal(a, c)                           # deterministic


def test_refuses_cpu_and_other_configs(cuda):
    from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res
    from future_urban_scene_generation_b200._lib import FusgError
    torch = cuda
    with pytest.raises(NotImplementedError):
        Vunet_fix_res(Namespace(up_mode='nearest', w_norm=False, drop_prob=0.0, vunet_256=False))
    m = Vunet_fix_res(_cfg()).eval()
    with pytest.raises(FusgError):
        m(torch.zeros(1, 3, 256, 256), torch.zeros(1, 6, 256, 256))
