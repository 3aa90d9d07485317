"""VUNet input packing (trajectory_inference.py:205-227) at BASELINE config 5 shape: V vehicles x T steps, 1920x1080 frames.
usage: python scripts/bench_pack.py [V=30] [T=20] [reps=5]   -> one JSON line: time of the public call (Python packing of
the per-item rectangles + their H2D + three launches) and the reference lines on the host CPU on a bounded sample."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from future_urban_scene_generation_b200 import _lib, synth
from future_urban_scene_generation_b200.frame_ops import pack_vunet_inputs_batch
from oracle import frame_oracle as FO

V = int(sys.argv[1]) if len(sys.argv) > 1 else 30
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
Hf, Wf = 1080, 1920
frames = np.random.default_rng(0).integers(0, 256, (T, Hf, Wf, 3), dtype=np.uint8)
full, rect_in, rects, fidx = [], [], [], []
for t in range(T):
    for v in range(V):
        m, ns, nd = synth.make_pack_case(2000 + t * V + v, (Hf, Wf))
        ys, xs = np.nonzero(~m)
        x0, y0, x1, y1 = xs.min(), ys.min(), xs.max(), ys.max()
        if len(full) < 2 * V:
            full.append((m, ns, nd))
        rect_in.append((np.ascontiguousarray(m[y0:y1 + 1, x0:x1 + 1]), np.ascontiguousarray(ns[y0:y1 + 1, x0:x1 + 1]), np.ascontiguousarray(nd[y0:y1 + 1, x0:x1 + 1])))
        rects.append((int(x0), int(y0), int(x1 - x0 + 1), int(y1 - y0 + 1)))
        fidx.append(t)
B = len(fidx)
d_frames = torch.from_numpy(frames).cuda()
args = (d_frames, fidx, [r[0] for r in rect_in], [r[1] for r in rect_in], [r[2] for r in rect_in])
x, y, bbox = pack_vunet_inputs_batch(*args, rects=rects)
torch.cuda.synchronize()
n0 = _lib.kernel_launches()
t0 = time.perf_counter()
for _ in range(reps):
    x, y, bbox = pack_vunet_inputs_batch(*args, rects=rects)
torch.cuda.synchronize()
call_ms = (time.perf_counter() - t0) * 1e3 / reps
launches = (_lib.kernel_launches() - n0) // reps
# CPU: reference lines on a bounded sample
try:
    import cv2
    rz, kind = (lambda a, ds: cv2.resize(a, ds)), "reference lines with cv2.resize"
except Exception:
    rz, kind = FO.resize_linear_u8, "numpy oracle"
n_cpu = len(full)
t0 = time.perf_counter()
outs = [FO.pack_vunet_inputs(frames[fidx[b]], *full[b], resize=rz) for b in range(n_cpu)]
cpu_ms_item = (time.perf_counter() - t0) * 1e3 / n_cpu
for b in range(n_cpu):
    assert np.array_equal(x[b].cpu().numpy(), outs[b][0]) and np.array_equal(y[b].cpu().numpy(), outs[b][1]), b
out_bytes = x.numel() * 4 + y.numel() * 4
print(json.dumps({"workload": f"VUNet input packing, {V} vehicles x {T} steps from {T} frames of {Wf}x{Hf}", "items": B,
                  "call_ms_incl_python_and_h2d_of_rect_inputs": call_ms, "items_per_s": B / call_ms * 1e3, "kernel_launches_per_call": launches,
                  "output_bytes": out_bytes, "cpu_ms_per_item": cpu_ms_item, "cpu_kind": kind, "cpu_sample_items": n_cpu,
                  "cpu_ms_whole_clip_extrapolated": cpu_ms_item * B, "checked": f"first {n_cpu} items bit-equal to the CPU lines"}))
