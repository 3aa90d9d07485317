"""The normal-sketch renderer alone at BASELINE config 5's shape: 30 vehicles x (20 moved poses + 1 source pose) on 1920x1080.
usage: python scripts/bench_render.py [reps]      (ncu: -k regex:k_render)"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from future_urban_scene_generation_b200 import synth
from future_urban_scene_generation_b200.warp_learn.render import render_normals_batch, MeshOnDevice

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
V, S, H, W = 30, 20, 1080, 1920
dev = torch.device("cuda")
meshes = [MeshOnDevice(*synth.make_car_mesh(v % 10)) for v in range(V)]
K = np.array([[1400.0, 0, W / 2], [0, 1400.0, H / 2], [0, 0, 1]])
E = np.zeros((V, 3, 4))
rng = np.random.default_rng(0)
for v in range(V):
    yaw = rng.uniform(0, 2 * math.pi)
    R = np.array([[math.cos(yaw), -math.sin(yaw), 0], [0, 0, -1], [math.sin(yaw), math.cos(yaw), 0]])      # car z up -> camera y down
    E[v, :, :3] = R
    E[v, :, 3] = [rng.uniform(-6, 6), 1.4, rng.uniform(14, 30)]
E4 = torch.from_numpy(E).to(dev)
K4 = torch.from_numpy(np.broadcast_to(K, (V, 3, 3)).copy()).to(dev)
th = rng.uniform(-0.3, 0.3, (V * S,))
rot = np.stack([np.array([[math.cos(t), -math.sin(t), 0], [math.sin(t), math.cos(t), 0], [0, 0, 1]]) for t in th])
tr = np.concatenate([rng.uniform(-2, 2, (V * S, 2)), np.zeros((V * S, 1))], 1)
rot_t, tr_t = torch.from_numpy(rot).to(dev), torch.from_numpy(tr).to(dev)
dst_n = torch.empty((V * S, H, W, 3), dtype=torch.uint8, device=dev)
dst_b = torch.empty((V * S, H, W), dtype=torch.bool, device=dev)
src_n = torch.empty((V, H, W, 3), dtype=torch.uint8, device=dev)
src_b = torch.empty((V, H, W), dtype=torch.bool, device=dev)


def run():
    for v in range(V):
        sl = slice(v * S, (v + 1) * S)
        render_normals_batch(meshes[v], E4[v:v + 1].expand(S, 3, 4), K4[v:v + 1].expand(S, 3, 3), H, W, rot=rot_t[sl], tr=tr_t[sl], out=(dst_n[sl], dst_b[sl]))
        render_normals_batch(meshes[v], E4[v:v + 1], K4[v:v + 1], H, W, out=(src_n[v:v + 1], src_b[v:v + 1]))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    run()
e1.record()
host_ms = (time.perf_counter() - t0) * 1e3 / reps
torch.cuda.synchronize()
print(f"{V * (S + 1)} renders at {W}x{H}: {e0.elapsed_time(e1) / reps:.2f} ms on the device per clip, {host_ms:.2f} ms of host time to enqueue; "
      f"object pixels per render: {float((~dst_b).sum()) / (V * S):.0f}")
