"""Plain PyTorch fp32 restatement of the reference VUNet forward -- TEST INFRASTRUCTURE ONLY.

Follows vunet/layers.py:6-221 and vunet/models.py:17-484 of the reference for the
`run_test.py:82-83` configuration (up_mode='subpixel', w_norm=True, drop_prob=0.2,
vunet_256=True, eval mode).  It is a *functional* restatement driven directly by a
state_dict with the reference's 336 keys, so it shares no code with the product's
module/engine.  Pinned by scripts/make_golden_vunet.py (run in the build container, where
/root/reference is importable): identical outputs to the reference module for identical
weights and CPU noise; the resulting fingerprints are committed in tests/golden/vunet_golden.json.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ layer primitives
def _weight(sd, path):
    """weight_norm(dim=0): w = g * v / ||v||_(1,2,3)   (layers.py:29-31)"""
    g, v = sd[path + ".conv.weight_g"], sd[path + ".conv.weight_v"]
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1))


def conv(sd, path, x, stride=1):
    w = _weight(sd, path)
    return F.conv2d(x, w, sd[path + ".conv.bias"], stride=stride, padding=w.shape[-1] // 2)


def nin(sd, path, x):                       # layers.py:42-58
    return conv(sd, path + ".layers.1", F.elu(x))


def residual(sd, path, x, skip=None):       # layers.py:83-105 (dropout is identity in eval)
    h = x if skip is None else torch.cat([x, skip], 1)
    return conv(sd, path + ".layers.2", F.elu(h)) + x


def space_to_depth(x):                      # layers.py:197-221, block 2, TF (block-major) order
    b, c, h, w = x.shape
    x = x.view(b, c, h // 2, 2, w // 2, 2)
    return x.permute(0, 3, 5, 1, 2, 4).reshape(b, 4 * c, h // 2, w // 2)


def depth_to_space(x):                      # layers.py:173-194
    b, c4, h, w = x.shape
    c = c4 // 4
    x = x.view(b, 2, 2, c, h, w)
    return x.permute(0, 3, 4, 1, 5, 2).reshape(b, c, 2 * h, 2 * w)


def upsample(sd, path, x):                  # layers.py:121-152, mode 'subpixel'
    return depth_to_space(conv(sd, path + ".depth4x", x))


def sampler(sd, path, x):                   # layers.py:158-170: CPU default generator, eval too
    mu = conv(sd, path + ".conv", x)
    return mu, mu + torch.randn(*mu.size()).to(mu.device) * 1.0


def down_block(sd, path, x):                # models.py:92-114
    x = conv(sd, path + ".down.down", x, stride=2)
    s0 = residual(sd, path + ".residual_0", x)
    s1 = residual(sd, path + ".residual_1", s0)
    return s1, [s0, s1]


def init_block(sd, path, x):                # models.py:141-163
    x = nin(sd, path + ".nin", x)
    s0 = residual(sd, path + ".residual_0", x)
    s1 = residual(sd, path + ".residual_1", s0)
    return s1, [s0, s1]


def up_block(sd, path, x, skip_a, skip_b):  # models.py:117-138
    x = residual(sd, path + ".residual_0", x, skip_a)
    x = residual(sd, path + ".residual_1", x, skip_b)
    return upsample(sd, path + ".up", x)


def ar_block(sd, path, x, skip_a, enc_mu=None):   # models.py:17-89
    x = residual(sd, path + ".residual_init", x, skip_a)
    x_ = space_to_depth(residual(sd, path + ".residual_s2d", x))
    g = None
    if enc_mu is not None:
        g = list(torch.split(space_to_depth(enc_mu), 128, 1))
        for k in range(3):
            g[k] = nin(sd, f"{path}.nin_{k}", g[k])
    mus, zs = [], []
    for k in range(4):
        mu_k, z_k = sampler(sd, f"{path}.sampler_{k}", x_)
        mus.append(mu_k)
        zs.append(z_k)
        if k < 3:
            skip = g[k] if g is not None else nin(sd, f"{path}.nin_{k}", z_k)
            x_ = residual(sd, f"{path}.residual_{k}", x_, skip)
    return x, depth_to_space(torch.cat(mus, 1)).contiguous(), depth_to_space(torch.cat(zs, 1))


# ------------------------------------------------------------------ the four sub-forwards
def forward_enc_up(sd, x):                  # models.py:333-353
    x, _ = init_block(sd, "app_encoder_1", x)
    for name in ("app_encoder_1_a", "app_encoder_1_b", "app_encoder_1_c", "app_encoder_2", "app_encoder_3"):
        x, _ = down_block(sd, name, x)
    skips = [nin(sd, "app_skip_3_c", x)]
    x, sl = down_block(sd, "app_encoder_4", x)
    outputs = [sl[-2], x]
    skips.append(nin(sd, "app_skip_4_c", x))
    return outputs, skips


def forward_enc_down(sd, outputs, skips):   # models.py:390-408
    x = conv(sd, "app_bottleneck", outputs[-1])
    x = residual(sd, "app_decoder_1_a", x, skips[-1])
    mu0, z0 = sampler(sd, "app_decoder_1_b", x)
    x_ = conv(sd, "app_decoder_1_c", torch.cat([outputs[-2], z0], 1))
    x = residual(sd, "app_decoder_1_d", x, x_)
    x = upsample(sd, "app_decoder_1_e", x)
    x = residual(sd, "app_decoder_2_a", x)
    mu1, z1 = sampler(sd, "app_decoder_2_b", x)
    return [mu0, mu1], [z0, z1]


def forward_dec_up(sd, y):                  # models.py:355-388
    skips = []
    x, sl = init_block(sd, "shape_encoder_1", y)
    skips += [nin(sd, "shape_skip_1_b", sl[-2]), nin(sd, "shape_skip_1_c", sl[-1])]
    for enc, sk in (("shape_encoder_1_a", "shape_skip_1_a"), ("shape_encoder_2", "shape_skip_2"),
                    ("shape_encoder_3", "shape_skip_3"), ("shape_encoder_4", "shape_skip_4"),
                    ("shape_encoder_5", "shape_skip_5"), ("shape_encoder_6", "shape_skip_6")):
        x, sl = down_block(sd, enc, x)
        skips += [nin(sd, sk + "_b", sl[-2]), nin(sd, sk + "_c", sl[-1])]
    return [x], skips


def forward_dec_down(sd, outputs, skips, enc_down_mu=()):   # models.py:410-459 (pops `skips`)
    mu, z = [], []
    x = conv(sd, "shape_bottleneck", outputs[-1])
    for n, blk in enumerate(("shape_decoder_1", "shape_decoder_2")):
        skip_a, skip_b = skips.pop(), skips.pop()
        x, mu_n, z_n = ar_block(sd, blk, x, skip_a, None if len(enc_down_mu) == 0 else enc_down_mu[n])
        mu.append(mu_n)
        z.append(z_n)
        x = nin(sd, blk + "_n", torch.cat([x, z_n], 1))
        x = residual(sd, blk + "_o", x, skip_b)
        x = upsample(sd, blk + "_p", x)
    for blk in ("shape_decoder_3", "shape_decoder_4", "shape_decoder_5", "shape_decoder_5_a"):
        skip_a, skip_b = skips.pop(), skips.pop()
        x = up_block(sd, blk, x, skip_a, skip_b)
    skip_a, skip_b = skips.pop(), skips.pop()
    x = residual(sd, "shape_decoder_6.residual_0", x, skip_a)
    x = residual(sd, "shape_decoder_6.residual_1", x, skip_b)
    x = conv(sd, "shape_decoder_6.conv", x)
    assert not skips
    return x, mu, z


def forward(sd, y_tilde, x=None, mean_mode="mean_appearance"):   # models.py:461-481
    assert y_tilde.shape[-1] == 256
    assert mean_mode in ["mean_appearance", "mean_shape"]
    if mean_mode == "mean_appearance":
        out_e, skips_e = forward_enc_up(sd, x)
        mu_app, z_app = forward_enc_down(sd, out_e, skips_e)
        out_d, skips_d = forward_dec_up(sd, y_tilde)
        x_tilde, mu_shape, _ = forward_dec_down(sd, out_d, skips_d, z_app)
        return x_tilde, mu_app, mu_shape
    out_d, skips_d = forward_dec_up(sd, y_tilde)
    return forward_dec_down(sd, out_d, skips_d)[0]


# ------------------------------------------------------------------ deterministic weights
def conv_registry():
    """[(path, cout, cin, k)] in the reference's registration order."""
    reg = []

    def C(path, cout, cin, k):
        reg.append((path, cout, cin, k))

    def res(path, cin, cout):
        C(path + ".layers.2", cout, cin, 3)

    def down(path, cin, cout):
        C(path + ".down.down", cout, cin, 3)
        res(path + ".residual_0", cout, cout)
        res(path + ".residual_1", cout, cout)

    def init(path, cin, cout):
        C(path + ".nin.layers.1", cout, cin, 1)
        res(path + ".residual_0", cout, cout)
        res(path + ".residual_1", cout, cout)

    def ninl(path, cin, cout):
        C(path + ".layers.1", cout, cin, 1)

    def ar(path):
        res(path + ".residual_init", 256, 128)
        for k in range(3):
            C(f"{path}.sampler_{k}.conv", 128, 512, 3)
            res(f"{path}.residual_{k}", 1024, 512)
        C(f"{path}.sampler_3.conv", 128, 512, 3)
        for k in range(3):
            ninl(f"{path}.nin_{k}", 128, 512)
        res(path + ".residual_s2d", 128, 128)

    def up(path, cin, cmid, cout):
        res(path + ".residual_0", cin, cmid)
        res(path + ".residual_1", cin, cmid)
        C(path + ".up.depth4x", 4 * cout, cmid, 3)

    init("app_encoder_1", 6, 128)
    for n in ("app_encoder_1_a", "app_encoder_1_b", "app_encoder_1_c", "app_encoder_2", "app_encoder_3", "app_encoder_4"):
        down(n, 128, 128)
    ninl("app_skip_3_c", 128, 128)
    ninl("app_skip_4_c", 128, 128)
    C("app_bottleneck", 128, 128, 1)
    res("app_decoder_1_a", 256, 128)
    C("app_decoder_1_b.conv", 128, 128, 3)
    C("app_decoder_1_c", 128, 256, 1)
    res("app_decoder_1_d", 256, 128)
    C("app_decoder_1_e.depth4x", 512, 128, 3)
    res("app_decoder_2_a", 128, 128)
    C("app_decoder_2_b.conv", 128, 128, 3)
    init("shape_encoder_1", 3, 32)
    down("shape_encoder_1_a", 32, 32)
    down("shape_encoder_2", 32, 64)
    down("shape_encoder_3", 64, 128)
    for n in ("shape_encoder_4", "shape_encoder_5", "shape_encoder_6"):
        down(n, 128, 128)
    for n, c in (("shape_skip_1", 32), ("shape_skip_1_a", 32), ("shape_skip_2", 64), ("shape_skip_3", 128),
                 ("shape_skip_4", 128), ("shape_skip_5", 128), ("shape_skip_6", 128)):
        ninl(n + "_b", c, c)
        ninl(n + "_c", c, c)
    C("shape_bottleneck", 128, 128, 1)
    for n in ("shape_decoder_1", "shape_decoder_2"):
        ar(n)
        ninl(n + "_n", 256, 128)
        res(n + "_o", 256, 128)
        C(n + "_p.depth4x", 512, 128, 3)
    up("shape_decoder_3", 256, 128, 128)
    up("shape_decoder_4", 256, 128, 64)
    up("shape_decoder_5", 128, 64, 32)
    up("shape_decoder_5_a", 64, 32, 32)
    res("shape_decoder_6.residual_0", 64, 32)
    res("shape_decoder_6.residual_1", 64, 32)
    C("shape_decoder_6.conv", 3, 32, 3)
    return reg


def make_state_dict(seed=0):
    """Deterministic random-init weights with the reference's key set/order/shapes.
    (Checkpoints are not available offline; values are this oracle's own draw, NOT the
    reference's init stream -- parity tests load the same dict into both sides.)"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for path, cout, cin, k in conv_registry():
        fan_in = cin * k * k
        bound = 1.0 / fan_in ** 0.5
        v = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) * bound
        sd[path + ".conv.bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound
        sd[path + ".conv.weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1, 1) * (0.8 + 0.4 * torch.rand((cout, 1, 1, 1), generator=g))
        sd[path + ".conv.weight_v"] = v
    return sd
