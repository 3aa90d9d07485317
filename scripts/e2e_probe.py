import sys, os, time
sys.path.insert(0, os.getcwd())
from argparse import Namespace
import numpy as np, torch
from future_urban_scene_generation_b200 import synth
from future_urban_scene_generation_b200.pipeline import NovelViewPipeline
from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res
B=64
torch.manual_seed(0)
m = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).cuda().eval()
wb = synth.make_warp_batch(0, B); xs, ys = synth.make_vunet_inputs(0, B)
host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in wb.items()}
host["x"], host["y"] = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ys).pin_memory()
def run(pipe, n):
    last=None
    for _ in range(n):
        t=pipe.submit(host)
        if last is not None: pipe.result(last)
        last=t
    pipe.result(last); torch.cuda.synchronize()
for depth, shared in ((2, False), (2, True), (3, False), (3, True), (1, True)):
    pipe = NovelViewPipeline(m, depth=depth, shared_stream=shared)
    run(pipe, depth + 3)
    T0=time.perf_counter(); run(pipe, 30); dt=(time.perf_counter()-T0)*1e3/30
    print(f"depth={depth} shared_stream={shared}: {dt:.2f} ms/step e2e")
    del pipe; torch.cuda.empty_cache()

# --- component costs of one e2e step -------------------------------------------------------------
dev = torch.device("cuda")
dst = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
s = torch.cuda.Stream()
def t_copy():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for k, v in host.items(): dst[k].copy_(v, non_blocking=True)
        e1.record(s)
    s.synchronize()
    return e0.elapsed_time(e1)
t_copy()
nb = sum(v.numel() * v.element_size() for v in host.values())
ms = min(t_copy() for _ in range(5))
print(f"H2D {nb/1e6:.1f} MB: {ms:.2f} ms ({nb/ms/1e6:.1f} GB/s)")
pipe = NovelViewPipeline(m, depth=2, shared_stream=True)
run(pipe, 4)
out = pipe.result(pipe.n - 1)
devout = {k: v.to(dev) for k, v in out.items()}
def t_d2h():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for k, v in devout.items(): out[k].copy_(v, non_blocking=True)
        e1.record(s)
    s.synchronize()
    return e0.elapsed_time(e1)
t_d2h()
nb2 = sum(v.numel() * v.element_size() for v in out.values())
ms2 = min(t_d2h() for _ in range(5))
print(f"D2H {nb2/1e6:.1f} MB: {ms2:.2f} ms ({nb2/ms2/1e6:.1f} GB/s)")
sl = pipe.slots[0]
T0 = time.perf_counter()
for _ in range(10): pipe._draw_noise(sl)
print(f"noise draw (host): {(time.perf_counter()-T0)*100:.3f} ms/step, {sum(x.numel() for x in sl.noise_stage)*4/1e6:.2f} MB")
T0 = time.perf_counter()
for _ in range(10):
    t = pipe.submit(host); pipe.result(t)
torch.cuda.synchronize()
print(f"unpipelined submit+result latency: {(time.perf_counter()-T0)*100:.2f} ms")
