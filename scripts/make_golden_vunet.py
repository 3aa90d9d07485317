"""Pins oracle/vunet_oracle.py against the imported reference (runs only where /root/reference
exists, i.e. in the build container) and writes tests/golden/vunet_golden.json:
  * the reference's state_dict key list hash and shapes == the oracle registry's,
  * reference module vs oracle outputs for identical weights + CPU noise (max-abs must be ~1e-5),
  * fingerprints (strided samples + means) of the oracle outputs for seeded weights/inputs,
    so the GPU box can re-check the oracle and the CUDA path without the reference.
"""
import hashlib
import json
import os
import sys
from argparse import Namespace

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")
import warnings
warnings.filterwarnings("ignore")
import numpy as np
import torch

from oracle import vunet_oracle as VO
from future_urban_scene_generation_b200 import synth
from vunet.models import Vunet_fix_res          # the reference

torch.set_num_threads(8)
ref = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).eval()
ref_sd = ref.state_dict()
sd = VO.make_state_dict(0)
assert list(ref_sd.keys()) == list(sd.keys()), "key order differs"
for k in sd:
    assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
key_sha1 = hashlib.sha1("\n".join(sd.keys()).encode()).hexdigest()
print("keys", len(sd), key_sha1)
ref.load_state_dict(sd, strict=True)

gold = {"key_sha1": key_sha1, "n_keys": len(sd), "n_params": int(sum(v.numel() for v in sd.values())), "cases": []}
for B, start in ((1, 0), (2, 5)):
    x, y = synth.make_vunet_inputs(start, B)
    x, y = torch.from_numpy(x), torch.from_numpy(y)
    with torch.no_grad():
        torch.manual_seed(1)
        r_x, r_mua, r_mus = ref(y, x)
        torch.manual_seed(1)
        o_x, o_mua, o_mus = VO.forward(sd, y, x)
        # traj_test style: sub-forwards with mu_app (trajectory_inference.py:230-233)
        torch.manual_seed(2)
        oe, se = ref.forward_enc_up(x); mu_app, z_app = ref.forward_enc_down(oe, se)
        od, sdn = ref.forward_dec_up(y); r_img, r_mu2, r_z2 = ref.forward_dec_down(od, sdn, mu_app)
        torch.manual_seed(2)
        oe, se = VO.forward_enc_up(sd, x); mu_app2, z_app2 = VO.forward_enc_down(sd, oe, se)
        od, sdn = VO.forward_dec_up(sd, y); o_img, o_mu2, o_z2 = VO.forward_dec_down(sd, od, sdn, mu_app2)
        torch.manual_seed(3)
        r_ms = ref(y, None, mean_mode='mean_shape')
        torch.manual_seed(3)
        o_ms = VO.forward(sd, y, None, mean_mode='mean_shape')
    errs = {
        "x_tilde": float((r_x - o_x).abs().max()),
        "mu_app": max(float((a - b).abs().max()) for a, b in zip(r_mua, o_mua)),
        "mu_shape": max(float((a - b).abs().max()) for a, b in zip(r_mus, o_mus)),
        "traj_img": float((r_img - o_img).abs().max()),
        "mean_shape": float((r_ms - o_ms).abs().max()),
    }
    print("B", B, "oracle-vs-reference max-abs", errs)
    assert max(errs.values()) < 5e-5, errs

    def fp(t):
        t = t.detach().float()
        flat = t.flatten()
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        return {"shape": list(t.shape), "mean": float(t.double().mean()), "absmean": float(t.double().abs().mean()),
                "samples": [float(v) for v in flat[idx]]}
    gold["cases"].append({
        "B": B, "start": start, "noise_seed": 1, "oracle_vs_reference_maxabs": errs,
        "x_tilde": fp(o_x), "mu_app0": fp(o_mua[0]), "mu_app1": fp(o_mua[1]), "mu_shape0": fp(o_mus[0]), "mu_shape1": fp(o_mus[1]),
        "traj_img_seed2": fp(o_img), "mean_shape_seed3": fp(o_ms),
        "x_tilde_range": [float(o_x.min()), float(o_x.max())],
    })
with open(os.path.join(ROOT, "tests", "golden", "vunet_golden.json"), "w") as f:
    json.dump(gold, f, indent=1)
print("wrote tests/golden/vunet_golden.json")
