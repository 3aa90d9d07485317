"""BASELINE config 5 shape for the paste-back row: V vehicles x T future steps pasted into T frames of 1920x1080.
usage: python scripts/bench_paste.py [V=30] [T=20] [reps=5]
Prints one JSON line: device time of fusg_paste_back for the whole clip (inputs resident), the same through host
buffers (H2D of masks/crops, D2H of frames), and the reference lines on the host CPU (cv2 when importable, else the
numpy oracle) on a bounded sample."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from future_urban_scene_generation_b200 import synth
from future_urban_scene_generation_b200.frame_ops import paste_back_batch, paste_back_packed, prepare_paste
from oracle import frame_oracle as FO

V = int(sys.argv[1]) if len(sys.argv) > 1 else 30
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
Hf, Wf = 1080, 1920
rng = np.random.default_rng(0)
frames = rng.integers(0, 256, (T, Hf, Wf, 3), dtype=np.uint8)
crops, masks_full, masks_rect, rects, infos, fidx = [], [], [], [], [], []
for t in range(T):
    for v in range(V):
        bbox, mask, net = synth.make_paste_case(1000 + t * V + v, (Hf, Wf))
        x0, y0, x1, y1 = bbox
        crops.append(net); infos.append(FO.square_crop_info((Hf, Wf), bbox)); fidx.append(t)
        masks_full.append(mask); masks_rect.append(np.ascontiguousarray(mask[y0:y1 + 1, x0:x1 + 1])); rects.append((x0, y0, x1 - x0 + 1, y1 - y0 + 1))
crops = np.stack(crops)
B = len(fidx)

# device-resident
d_frames = torch.from_numpy(frames).cuda()
d_crops = torch.from_numpy(crops).cuda()
plan = prepare_paste(masks_rect, infos, fidx, (Hf, Wf), T, mask_rects=rects)
out = paste_back_packed(d_frames.clone(), d_crops, plan)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
work = [d_frames.clone() for _ in range(reps)]
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record()
for r in range(reps):
    paste_back_packed(work[r], d_crops, plan)
e1.record()
torch.cuda.synchronize()
wall_ms = (time.perf_counter() - t0) * 1e3 / reps
dev_ms = e0.elapsed_time(e1) / reps

# end to end: masks + crops from pinned host memory, frames back to the host
t0 = time.perf_counter()
for r in range(reps):
    o = paste_back_batch(torch.from_numpy(frames).cuda(), crops, masks_rect, infos, fidx, mask_rects=rects)
    host = o.cpu()
e2e_ms = (time.perf_counter() - t0) * 1e3 / reps

# CPU: the reference lines, sequential, on a bounded sample of the clip
try:
    import cv2
    rz, kind = (lambda a, ds: cv2.resize(a, ds)), "reference lines with cv2.resize"
except Exception:
    rz, kind = FO.resize_linear_u8, "numpy oracle"
n_cpu = min(B, 2 * V)
ref = frames[:2].copy()
t0 = time.perf_counter()
for b in range(n_cpu):
    FO.paste_back(ref[fidx[b]], crops[b], infos[b], masks_full[b], resize=rz)
cpu_ms_item = (time.perf_counter() - t0) * 1e3 / n_cpu
assert np.array_equal(out[:2].cpu().numpy(), ref), "GPU paste-back differs from the CPU reference lines"

mask_bytes = sum(m.size for m in masks_rect)
masked_px = sum(int(m.sum()) for m in masks_rect)
alg = mask_bytes + crops.nbytes + 2 * T * Hf * Wf * 4 + masked_px * 3     # masks + crops + owner (clear, read) + pasted pixels
print(json.dumps({"workload": f"paste-back, {V} vehicles x {T} steps into {T} frames of {Wf}x{Hf} (config 5)", "items": B,
                  "device_ms": dev_ms, "items_per_s": B / dev_ms * 1e3, "host_wall_ms_incl_python": wall_ms,
                  "e2e_ms_host_buffers": e2e_ms, "algorithmic_bytes": alg, "achieved_GBps": alg / dev_ms / 1e6,
                  "cpu_ms_per_item": cpu_ms_item, "cpu_kind": kind, "cpu_sample_items": n_cpu,
                  "cpu_ms_whole_clip_extrapolated": cpu_ms_item * B, "checked": "first 2 frames bit-equal to the CPU lines"}))
