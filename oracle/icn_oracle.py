"""Plain PyTorch fp32 restatement of the reference ICN generator forward -- TEST INFRASTRUCTURE ONLY.

Follows warp_learn/models.py:15-208 of the reference for the configuration run_test.py:75 builds
(`G_Resnet(21)`: output_nc=3, num_downs=2, n_res=3, ngf=64, norm='inst', nl_layer='relu',
pad_type='reflect').  It is a *functional* restatement driven directly by a state_dict with the
reference's 40 keys, so it shares no code with the product's module/engine.  Pinned by
scripts/make_golden_icn.py (run in the build container, where /root/reference is importable):
identical outputs to the reference module for identical weights; fingerprints are committed in
tests/golden/icn_golden.json.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def registry(input_nc=21, output_nc=3, ngf=64, n_down=2, n_res=3):
    """(state_dict prefix, kind, cout, cin, k) in the reference's registration order; kind 'ln' blocks register
    norm.gamma / norm.beta BEFORE conv.weight / conv.bias (Conv2dBlock builds its norm first, models.py:51-81)."""
    reg = [("enc_content.model.0", "conv", ngf, input_nc, 7)]
    dim = ngf
    for i in range(n_down):
        reg.append((f"enc_content.model.{1 + i}", "conv", 2 * dim, dim, 4))
        dim *= 2
    for r in range(n_res):
        for j in range(2):
            reg.append((f"enc_content.model.{1 + n_down}.model.{r}.model.{j}", "conv", dim, dim, 3))
    for r in range(n_res):
        for j in range(2):
            reg.append((f"dec.model.0.model.{r}.model.{j}", "conv", dim, dim, 3))
    for i in range(n_down):
        reg.append((f"dec.model.{2 + 2 * i}", "ln", dim // 2, dim, 5))
        dim //= 2
    reg.append((f"dec.model.{1 + 2 * n_down}", "conv", output_nc, dim, 7))
    return reg


def make_state_dict(seed=0, **kw):
    """Deterministic random weights with the reference's key set / order / shapes (this oracle's own draw, not the
    reference's init stream: parity tests load the same dict into both sides)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for prefix, kind, cout, cin, k in registry(**kw):
        bound = 1.0 / (cin * k * k) ** 0.5
        if kind == "ln":
            sd[prefix + ".norm.gamma"] = torch.rand((cout,), generator=g)
            sd[prefix + ".norm.beta"] = (torch.rand((cout,), generator=g) - 0.5) * 0.2
        sd[prefix + ".conv.weight"] = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) * bound
        sd[prefix + ".conv.bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound
    return sd


def conv_block(sd, prefix, x, stride, pad, norm, act):
    """Conv2dBlock (models.py:38-91): ReflectionPad2d -> Conv2d(bias) -> norm -> activation."""
    x = F.conv2d(F.pad(x, (pad, pad, pad, pad), mode="reflect"), sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=stride)
    if norm == "inst":                                # nn.InstanceNorm2d(track_running_stats=False): biased variance, no affine
        mean = x.mean(dim=(2, 3), keepdim=True)
        var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
        x = (x - mean) / torch.sqrt(var + EPS)
    elif norm == "ln":                                # the reference's own LayerNorm (models.py:15-35): unbiased std, (std + eps)
        flat = x.reshape(x.shape[0], -1)
        mean = flat.mean(1).view(-1, 1, 1, 1)
        std = flat.std(1).view(-1, 1, 1, 1)
        x = (x - mean) / (std + EPS)
        x = x * sd[prefix + ".norm.gamma"].view(1, -1, 1, 1) + sd[prefix + ".norm.beta"].view(1, -1, 1, 1)
    if act == "relu":
        x = F.relu(x)
    elif act == "tanh":
        x = torch.tanh(x)
    return x


def res_blocks(sd, prefix, x, n_res=3):
    """ResBlocks / ResBlock (models.py:94-125): x + block(act=none)(block(relu)(x))."""
    for r in range(n_res):
        h = conv_block(sd, f"{prefix}.model.{r}.model.0", x, 1, 1, "inst", "relu")
        x = x + conv_block(sd, f"{prefix}.model.{r}.model.1", h, 1, 1, "inst", "none")
    return x


def encode(sd, image, n_down=2, n_res=3):
    """ContentEncoder (models.py:128-149)."""
    x = conv_block(sd, "enc_content.model.0", image, 1, 3, "inst", "relu")
    for i in range(n_down):
        x = conv_block(sd, f"enc_content.model.{1 + i}", x, 2, 1, "inst", "relu")
    return res_blocks(sd, f"enc_content.model.{1 + n_down}", x, n_res)


def decode(sd, content, n_up=2, n_res=3):
    """Decoder (models.py:164-188): ResBlocks, then n_up x [nearest 2x upsample, 5x5 block with LayerNorm], then 7x7 tanh."""
    x = res_blocks(sd, "dec.model.0", content, n_res)
    for i in range(n_up):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = conv_block(sd, f"dec.model.{2 + 2 * i}", x, 1, 2, "ln", "relu")
    return conv_block(sd, f"dec.model.{1 + 2 * n_up}", x, 1, 3, "none", "tanh")


def forward(sd, image):
    """G_Resnet.forward (models.py:204-207)."""
    return decode(sd, encode(sd, image))


def flops_per_crop(res=256, input_nc=21):
    """Algorithmic FLOPs (2*MACs, real channels) of one forward at res x res."""
    total, h = 0, res
    for prefix, kind, cout, cin, k in registry(input_nc=input_nc):
        if prefix in ("enc_content.model.1", "enc_content.model.2"):
            h //= 2
        if kind == "ln":
            h *= 2
        total += 2 * h * h * cout * cin * k * k
    return total
