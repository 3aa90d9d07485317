"""Resident vs end-to-end step time of NovelViewPipeline across slot depth / stream settings (B = 64).
usage: python scripts/e2e_probe.py"""
import sys, os, time
sys.path.insert(0, os.getcwd())
from argparse import Namespace
from collections import deque
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
def run(pipe, n, resident=False, lag=None):
    lag = lag or pipe.depth - 1
    pend = deque()
    for _ in range(n):
        pend.append(pipe.submit(host, resident=resident))
        if len(pend) > lag:
            t = pend.popleft()
            pipe.wait(t) if resident else pipe.result(t)
    while pend:
        t = pend.popleft()
        pipe.wait(t) if resident else pipe.result(t)
    torch.cuda.synchronize()
for depth, shared in ((2, False), (2, True), (3, True), (3, False)):
    pipe = NovelViewPipeline(m, depth=depth, shared_stream=shared)
    run(pipe, depth + 3)
    T0=time.perf_counter(); run(pipe, 40, resident=True); dt=(time.perf_counter()-T0)*1e3/40
    T0=time.perf_counter(); run(pipe, 40); dt2=(time.perf_counter()-T0)*1e3/40
    print(f"depth={depth} shared_stream={shared}: resident {dt:.2f} ms/step, e2e {dt2:.2f} ms/step")
    del pipe; torch.cuda.empty_cache()
