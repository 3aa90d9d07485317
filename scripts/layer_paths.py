"""Launch-ordered list of the conv layers of one VUNet forward (single stream), to join with an ncu launch list.
usage: ncu --metrics gpu__time_duration.sum -k regex:k_conv --csv --log-file L.csv python scripts/layer_paths.py 64 > paths.txt
The LAST len(paths) k_conv_* launches of L.csv are the layers of paths.txt in order."""
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
m.fork_branches = False
e = m.engine()
x, y = synth.make_vunet_inputs(0, B)
x, y = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
e.noise_provider = lambda b, c, h, w: torch.zeros((b, h, w, c), device="cuda")
m(y, x)
torch.cuda.synchronize()
e.profile = []
m(y, x)
torch.cuda.synchronize()
for p, impl, fl, e0, e1 in e.profile:
    print(p, impl, fl)
