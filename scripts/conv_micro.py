"""Micro-benchmark of single VUNet layers through the engine (for ncu captures).
usage: python scripts/conv_micro.py <layer> [B] [reps]
  layer: nin6 | res128 | res128raw | res32 | cat64 | down128 | skip32 | out3 | down128only | up32"""
import os
import sys
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res

layer = sys.argv[1] if len(sys.argv) > 1 else "res128"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
m = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).cuda().eval()
e = m.engine()


def rnd(c, h):
    a = e._act(B, c, h, h)
    a.raw.normal_()
    a.elu.normal_()
    return a


if layer == "nin6":
    x = rnd(e._w["app_encoder_1.nin.layers.1"][4], 256)
    fn = lambda: e.init_block("app_encoder_1", x, B)          # nin + 2 residuals
    fn = lambda: e.conv("app_encoder_1.nin.layers.1", [(x, "elu")], outs=e._plain_outs(e._act(B, 128, 256, 256)), B=B)
    flops = 2.0 * B * 65536 * 128 * 6
elif layer == "res128":
    x = rnd(128, 256)
    fn = lambda: e.residual("app_encoder_1.residual_0", x, B=B)
    flops = 2.0 * B * 65536 * 128 * 128 * 9
elif layer == "res128raw":
    x = rnd(128, 256)
    fn = lambda: e.residual("app_encoder_1.residual_1", x, B=B, elu=False)
    flops = 2.0 * B * 65536 * 128 * 128 * 9
elif layer == "res32":
    x = rnd(32, 256)
    fn = lambda: e.residual("shape_encoder_1.residual_0", x, B=B)
    flops = 2.0 * B * 65536 * 32 * 32 * 9
elif layer == "cat64":
    x, s = rnd(32, 256), rnd(32, 256)
    fn = lambda: e.residual("shape_decoder_6.residual_0", x, s, B=B)
    flops = 2.0 * B * 65536 * 64 * 32 * 9
elif layer == "skip32":
    x = rnd(32, 256)
    fn = lambda: e.nin("shape_skip_1_b", [x], B, raw=True)
    flops = 2.0 * B * 65536 * 32 * 32
elif layer == "skip32e":
    x = rnd(32, 256)
    fn = lambda: e.nin("shape_skip_1_b", [x], B, raw=False)
    flops = 2.0 * B * 65536 * 32 * 32
elif layer == "out3":
    x = rnd(32, 256)
    xt = e._empty(B, 3, 256, 256, dtype=torch.float32)
    from future_urban_scene_generation_b200.vunet.engine import OutSpec
    fn = lambda: e.conv("shape_decoder_6.conv", [(x, "raw")], outs=[OutSpec(layout=1, tensor=xt)], B=B)
    flops = 2.0 * B * 65536 * 32 * 3 * 9
elif layer == "down128only":
    x = rnd(128, 256)
    d = e._act(B, 128, 128, 128)
    fn = lambda: e.conv("app_encoder_1_a.down.down", [(x, "raw")], stride=2, outs=e._plain_outs(d), B=B)
    flops = 2.0 * B * 16384 * 128 * 128 * 9
elif layer == "up32":
    x = rnd(64, 128)
    fn = lambda: e.upsample("shape_decoder_5_a.up", x, B)
    flops = 2.0 * B * 16384 * 64 * 128 * 9
elif layer == "down128":
    x = rnd(128, 256)
    fn = lambda: e.down_block("app_encoder_1_a", x, B)
    flops = 0
for _ in range(2):
    fn()
torch.cuda.synchronize()
from future_urban_scene_generation_b200 import _lib
print("plan [msub, pair, halo, ksplit, stages, group, w_resident, fast_epi] =", _lib.conv_last_plan())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{layer} B={B}: {ms:.3f} ms/launch, {flops / ms / 1e9:.1f} TFLOP/s")
