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
pipe = NovelViewPipeline(m, depth=2)
last=None
for _ in range(4):
    t=pipe.submit(host)
    if last is not None: pipe.result(last)
    last=t
pipe.result(last)
print("threads", torch.get_num_threads())
t0=time.perf_counter()
for _ in range(10): pipe._draw_noise(pipe.slots[0])
torch.cuda.synchronize(); print("draw_noise ms", (time.perf_counter()-t0)*100)
ts=[];tr=[]
last=None
T0=time.perf_counter()
for _ in range(20):
    a=time.perf_counter(); t=pipe.submit(host); b=time.perf_counter()
    if last is not None: pipe.result(last)
    c=time.perf_counter(); ts.append(b-a); tr.append(c-b); last=t
pipe.result(last); torch.cuda.synchronize()
print("per step ms", (time.perf_counter()-T0)*50, "submit cpu ms", 1e3*np.mean(ts), "result wait ms", 1e3*np.mean(tr))
# H2D alone
torch.cuda.synchronize(); a=time.perf_counter()
for _ in range(10):
    for k,v in host.items(): pipe.slots[0].inp[k].copy_(v, non_blocking=True)
torch.cuda.synchronize(); print("H2D inputs ms", (time.perf_counter()-a)*100)
# variants: which part of the e2e step slows the GPU?
def loop(n, keys=(), noise=False, late_wait=False):
    T0=time.perf_counter()
    for _ in range(n):
        slot = pipe.slots[pipe.n % 2]
        if slot.done is not None: slot.done.synchronize()
        with torch.cuda.stream(pipe.copy_stream):
            for k in keys: slot.inp[k].copy_(host[k], non_blocking=True)
            if noise: pipe._draw_noise(slot)
            ev = torch.cuda.Event(); ev.record(pipe.copy_stream)
        slot.stream.wait_event(ev)
        with torch.cuda.stream(slot.stream):
            slot.graph.replay()
            comp = torch.cuda.Event(); comp.record(slot.stream)
        slot.done = comp
        pipe.n += 1
    torch.cuda.synchronize()
    return (time.perf_counter()-T0)*1e3/n
allk=tuple(host.keys())
for name,keys,noise in (("none",(),False),("noise only",(),True),("x only",("x",),False),("y only",("y",),False),("all inputs",allk,False),("all+noise",allk,True)):
    loop(4,keys,noise); print(name,"ms/step %.2f"%loop(20,keys,noise))
