"""Resident CUDA-graph replay time of one step (64 crops): quick A/B of kernel variants through env switches.
usage: [FUSG_...=..] python scripts/step_time.py [steps]"""
import os
import sys
import time
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from future_urban_scene_generation_b200 import synth
from future_urban_scene_generation_b200.pipeline import NovelViewPipeline
from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res

B = 64
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.manual_seed(0)
m = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).cuda().eval()
wb = synth.make_warp_batch(0, B)
xs, ys = synth.make_vunet_inputs(0, B)
host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in wb.items()}
host["x"], host["y"] = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ys).pin_memory()
for shared in (True, False):
    pipe = NovelViewPipeline(m, depth=2, shared_stream=shared)
    for _ in range(4):
        t = pipe.submit(host)
    pipe.result(t)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        T0 = time.perf_counter()
        for _ in range(n):
            t = pipe.submit(host, resident=True)
        pipe.wait(t)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - T0) * 1e3 / n)
    print(f"shared_stream={shared}: {best:.3f} ms/step resident")
    del pipe
