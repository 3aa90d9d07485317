"""One ICN generator forward for an ncu pass; prints the launch-ordered names of its 70 kernels.
usage: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \\
           --clock-control none -k regex:^k_ --csv --log-file L.csv python scripts/icn_ncu.py 64 > names.txt
       python scripts/icn_ncu_join.py names.txt L.csv > profiles/r1_icn_ncu.txt
The LAST len(names) k_* launches of L.csv are the launches of names.txt in order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from future_urban_scene_generation_b200 import synth
from future_urban_scene_generation_b200.warp_learn.models import G_Resnet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = G_Resnet(21).cuda().eval()
x = torch.from_numpy(synth.make_icn_inputs(0, min(B, 8), 256)).cuda()
x = x.repeat((B + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:B].contiguous()
e = m.engine()
e.profile = []
m(x)
torch.cuda.synchronize()
for name, kind, amount, _, _ in e.profile:
    print(name, kind, amount)
