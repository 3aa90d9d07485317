"""Per-launch CUDA-event timing of one VUNet forward: which layers eat the step.
usage: python scripts/profile_layers.py [B] > gpurun_out/layers.txt"""
import os
import sys
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from future_urban_scene_generation_b200 import synth
from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).cuda().eval()
e = m.engine()
m.fork_branches = False   # per-launch timing on one stream
x, y = synth.make_vunet_inputs(0, B)
x, y = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
bank = {}


def noise(b, c, h, w):
    k = (b, c, h, w)
    if k not in bank:
        bank[k] = torch.randn((b, h, w, c), device="cuda")
    return bank[k]


e.noise_provider = noise
for _ in range(3):
    m(y, x)
torch.cuda.synchronize()
e.profile = []
m(y, x)
torch.cuda.synchronize()
rows = [(p, impl, fl, e0.elapsed_time(e1)) for p, impl, fl, e0, e1 in e.profile]
tot = sum(r[3] for r in rows)
print(f"B={B} total conv time {tot:.3f} ms, {sum(r[2] for r in rows) / tot / 1e9:.1f} TFLOP/s")
print(f"{'layer':48s} impl {'GFLOP':>9s} {'ms':>8s} {'TFLOP/s':>8s} {'%':>6s}")
for p, impl, fl, ms in rows:
    print(f"{p:48s} {impl:4d} {fl / 1e9:9.2f} {ms:8.3f} {fl / ms / 1e9:8.1f} {100 * ms / tot:6.2f}")
