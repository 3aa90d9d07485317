"""ICN generator on one B200: crops/s and conv TFLOP/s of G_Resnet(21) at 256x256, random-init weights.
  python scripts/bench_icn.py [--crops 64] [--steps 10] [--out gpurun_out/icn.json]
Times whole forwards with CUDA events (inputs resident on the device), then one profiled forward per launch."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from future_urban_scene_generation_b200 import synth, _lib
from future_urban_scene_generation_b200.warp_learn.models import G_Resnet

ap = argparse.ArgumentParser()
ap.add_argument("--crops", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--out", default=None)
ap.add_argument("--layers", default=None, help="write the per-launch table here")
args = ap.parse_args()

torch.manual_seed(0)
m = G_Resnet(21).cuda().eval()
x = torch.from_numpy(synth.make_icn_inputs(0, min(args.crops, 8), 256)).cuda()
x = x.repeat((args.crops + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:args.crops].contiguous()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
for _ in range(args.warmup):
    y = m(x)
torch.cuda.synchronize()
n0 = _lib.kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    y = m(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
launches = (_lib.kernel_launches() - n0) // args.steps
eng = m.engine()
# per-launch profile (the stream is parked first so that events bracket kernels, not host gaps)
eng.profile = []
torch.cuda._sleep(int(40e-3 * 1.9e9))
m(x)
torch.cuda.synchronize()
rows = [(n, k, a, s.elapsed_time(e)) for n, k, a, s, e in eng.profile]
eng.profile = None
conv_ms = sum(t for _, k, _, t in rows if k == "flops")
conv_fl = sum(a for _, k, a, _ in rows if k == "flops")
oth_ms = sum(t for _, k, _, t in rows if k == "bytes")
oth_by = sum(a for _, k, a, _ in rows if k == "bytes")
lines = ["%-44s %8s %10s %8s" % ("launch", "ms", "TFLOP/s", "GB/s")]
for n, k, a, t in rows:
    lines.append("%-44s %8.3f %10s %8s" % (n, t, "%.1f" % (a / t / 1e9) if k == "flops" else "", "%.0f" % (a / t / 1e6) if k == "bytes" else ""))
table = "\n".join(lines)
if args.layers:
    open(args.layers, "w").write("B=%d forward %.3f ms; convs %.3f ms (%.1f TFLOP/s), norm/layout passes %.3f ms (%.0f GB/s)\n%s\n"
                                 % (args.crops, ms, conv_ms, conv_fl / conv_ms / 1e9, oth_ms, oth_by / oth_ms / 1e6, table))
res = {"metric": "ICN generator crops/s (G_Resnet(21) fp16 forward, 256x256)", "value": args.crops / ms * 1e3, "unit": "crops/s",
       "crops": args.crops, "ms_per_forward": ms, "steps": args.steps, "gpu_launches_per_forward": launches,
       "flops_per_crop": conv_fl / args.crops,
       "tflops_whole_forward": conv_fl / ms / 1e9,
       "roofline": {"bound": "tensor", "kernel": "k_conv_tc (18 bordered convolutions)", "achieved": conv_fl / conv_ms / 1e9,
                    "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": conv_fl / conv_ms / 1e9 / peaks["bf16_tflops_sustained"],
                    "conv_ms": conv_ms},
       "norm_passes": {"bound": "hbm", "ms": oth_ms, "achieved": oth_by / oth_ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": oth_by / oth_ms / 1e6 / peaks["hbm_gbs"]},
       "data": "synthetic", "dtype": "fp16 operands, fp32 accumulate"}
# get_icn_inputs on the device for the same batch (256 x 256 frames: crop + cv2.resize + 8-bit Lab + normalise, 7 images per crop)
from future_urban_scene_generation_b200.frame_ops import get_icn_inputs_batch
import numpy as np
pk = [synth.make_icn_pack_case(i, (256, 256)) for i in range(8)]
rep = (args.crops + 7) // 8
pl = torch.from_numpy(np.stack([c[0] for c in pk])).cuda().repeat(rep, 1, 1, 1, 1)[:args.crops].contiguous()
nm = torch.from_numpy(np.stack([c[1] for c in pk])).cuda().repeat(rep, 1, 1, 1)[:args.crops].contiguous()
mk = torch.from_numpy(np.stack([c[2] for c in pk])).cuda().repeat(rep, 1, 1)[:args.crops].contiguous()
ct = torch.from_numpy(np.stack([c[3] for c in pk])).cuda().repeat(rep, 1, 1, 1)[:args.crops].contiguous()
for _ in range(2):
    get_icn_inputs_batch(pl, nm, mk, ct)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    gen_in, _ = get_icn_inputs_batch(pl, nm, mk, ct)
e1.record()
torch.cuda.synchronize()
res["input_packing"] = {"ms_per_batch": e0.elapsed_time(e1) / 5, "note": "get_icn_inputs_batch through the public call (bbox kernel, bbox D2H for crop_info, "
                        "pack kernel), 256x256 frames; bit-exact vs the reference function"}
print(json.dumps(res))
if args.out:
    open(args.out, "w").write(json.dumps(res) + "\n")
