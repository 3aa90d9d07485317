"""Pins oracle/icn_oracle.py against the imported reference ICN generator (runs only where /root/reference exists,
i.e. in the build container) and writes tests/golden/icn_golden.json:
  * the reference's state_dict key list / order / shapes == the oracle registry's,
  * reference module vs oracle outputs for identical weights (max-abs must be ~1e-5),
  * fingerprints (strided samples + means) of the oracle outputs for seeded weights / inputs, so the GPU box can
    re-check the oracle and the CUDA path without the reference.
"""
import hashlib
import json
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")
import warnings
warnings.filterwarnings("ignore")
import torch

from oracle import icn_oracle as IO
from future_urban_scene_generation_b200 import synth
from warp_learn.models import G_Resnet          # the reference

torch.set_num_threads(8)
ref = G_Resnet(21).eval()                       # run_test.py:75
ref_sd = ref.state_dict()
sd = IO.make_state_dict(0)
assert list(ref_sd.keys()) == list(sd.keys()), "key order differs"
for k in sd:
    assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
key_sha1 = hashlib.sha1("\n".join(sd.keys()).encode()).hexdigest()
print("keys", len(sd), key_sha1)
ref.load_state_dict(sd, strict=True)


def fp(t):
    t = t.detach().float()
    flat = t.flatten()
    idx = torch.linspace(0, flat.numel() - 1, 64).long()
    return {"shape": list(t.shape), "mean": float(t.double().mean()), "absmean": float(t.double().abs().mean()),
            "samples": [float(v) for v in flat[idx]]}


gold = {"key_sha1": key_sha1, "n_keys": len(sd), "n_params": int(sum(v.numel() for v in sd.values())),
        "flops_per_crop_256": IO.flops_per_crop(256), "cases": []}
for B, start, res in ((1, 0, 256), (2, 3, 64), (1, 9, 128)):
    x = torch.from_numpy(synth.make_icn_inputs(start, B, res))
    with torch.no_grad():
        r = ref(x)
        o = IO.forward(sd, x)
        r_c = ref.enc_content(x)
        o_c = IO.encode(sd, x)
        r_d = ref.decode(r_c)
    errs = {"out": float((r - o).abs().max()), "content": float((r_c - o_c).abs().max()), "decode": float((r_d - o).abs().max())}
    print("B", B, "res", res, "oracle-vs-reference max-abs", errs)
    assert max(errs.values()) < 5e-5, errs
    gold["cases"].append({"B": B, "start": start, "res": res, "oracle_vs_reference_maxabs": errs, "out": fp(o), "content": fp(o_c),
                          "out_range": [float(o.min()), float(o.max())]})
with open(os.path.join(ROOT, "tests", "golden", "icn_golden.json"), "w") as f:
    json.dump(gold, f, indent=1)
print("wrote tests/golden/icn_golden.json")
