#!/usr/bin/env python
"""bench.py -- novel-view completion (fused planar warp + VUNet bf16 forward) throughput on B200.

Metric (BASELINE.json): vehicle crops/s through warp + VUNet at N GPUs, plus the roofline fraction of
the dominant kernel, next to the reference algorithm's CPU path timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU under torchrun)
  python bench.py --impl reference ...                           CPU arm: the oracle port of the reference path

One "step" = one pass of the hot path over one batch of synthetic crops (BASELINE config 2: 64 crops
per GPU): fusg_warp_fused on the batch, Vunet_fix_res.forward on the batch, to_image, and -- when
N > 1 -- the NCCL all-gather of the completed uint8 crops (the only collective; crops are sharded
contiguously by rank, weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vehicle crops/s novel-view completion (warp+VUNet)"
UNIT = "crops/s"
WARP_BYTES_PER_CROP = 256 * 256 * 3 * 6            # read the crop once + write 5 planes
VUNET_FLOPS_PER_CROP = 76_271_321_088              # 112 convs, 2*MACs


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the busy samples (idle samples before the first launch sit at low clocks)
        busy = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (test infrastructure used as the timed baseline)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, crops_per_step=1):
    """Times the reference algorithm's CPU path: oracle/warp_oracle.c (scalar C, 1 thread) for
    visibility x2 + homographies + masked warp, and oracle/vunet_oracle.py (torch fp32, all host
    threads) for the VUNet forward -- BASELINE config 1 (batch 1, --device cpu)."""
    import numpy as np
    import torch
    from future_urban_scene_generation_b200 import synth
    from oracle import warp_oracle as WO, vunet_oracle as VO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = VO.make_state_dict(0)
    t_warp, t_vunet = [], []
    for s in range(warmup + steps):
        tw = tv = 0.0
        for c in range(crops_per_step):
            idx = s * crops_per_step + c
            p = synth.make_pose_pair(idx)
            img = synth.make_crop(idx)
            x, y = synth.make_vunet_inputs(idx, 1)
            x, y = torch.from_numpy(x), torch.from_numpy(y)
            t0 = time.perf_counter()
            WO.warp_fused(img, p["src_kp"], p["dst_kp"], p["K"], p["E_src"], p["E_dst"], p["kp3d"])
            t1 = time.perf_counter()
            with torch.no_grad():
                VO.forward(sd, y, x)
            t2 = time.perf_counter()
            tw += t1 - t0
            tv += t2 - t1
        if s >= warmup:
            t_warp.append(tw)
            t_vunet.append(tv)
    total = sum(t_warp) + sum(t_vunet)
    n = steps * crops_per_step
    return {"value": n / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} crops, batch 1 (BASELINE config 1): C oracle warp {1e3 * sum(t_warp) / n:.1f} ms/crop (1 thread) + "
                      f"torch fp32 VUNet oracle {sum(t_vunet) / n:.3f} s/crop ({torch.get_num_threads()} threads)",
            "ms_per_step": 1e3 * total / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, max(1, args.warmup))
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "1 synthetic 256x256 vehicle crop per step: warp (visibility x2, homographies, masked warp) + VUNet fp32 forward, CPU"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from argparse import Namespace
    import numpy as np
    import torch
    import torch.distributed as dist
    from future_urban_scene_generation_b200 import synth, _lib
    from future_urban_scene_generation_b200.warp_learn import warp_batch
    from future_urban_scene_generation_b200.warp_learn.planes_utils import to_image_batch
    from future_urban_scene_generation_b200.vunet.models import Vunet_fix_res
    from future_urban_scene_generation_b200.parallel import shard_range, gather_crops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version/info lines must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    B = args.crops_per_rank
    peaks = _peaks()

    # ---- model: random-init weights of the reference architecture (checkpoints are not available offline)
    torch.manual_seed(0)
    model = Vunet_fix_res(Namespace(up_mode='subpixel', w_norm=True, drop_prob=0.2, vunet_256=True)).to(dev).eval()
    eng = model.engine()

    # ---- synthetic inputs for this rank's contiguous shard of crops (SURVEY.md §8d generators)
    first, last = shard_range(world * B, rank, world)
    assert last - first == B
    wb = synth.make_warp_batch(first, B)
    xs, ys = synth.make_vunet_inputs(first, B)
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in wb.items()}
    host["x"] = torch.from_numpy(xs).pin_memory()
    host["y"] = torch.from_numpy(ys).pin_memory()
    devin = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    from future_urban_scene_generation_b200.pipeline import NovelViewPipeline
    # The public batched API: NovelViewPipeline (CUDA-graph replay of the ~125 launches of a step, two
    # steps in flight).  N > 1 adds the NCCL all-gather of the completed crops (eager launches then).
    pipe = NovelViewPipeline(model, depth=2, gather_fn=(lambda c: gather_crops(c, world * B)) if world > 1 else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def run_steps(n, resident, pipe=pipe, host=host):
        # keep depth-1 steps in flight behind the one being submitted; every step's outputs are read on the host
        pending, last = collections.deque(), None
        for _ in range(n):
            last = pipe.submit(host, resident=resident)
            pending.append(last)
            if len(pending) >= pipe.depth and not resident:
                pipe.result(pending.popleft())         # an earlier step's outputs are on the host
        if resident:
            pipe.wait(last)
        else:
            while pending:
                pipe.result(pending.popleft())
        return last

    # warm-up: builds the slots (eager pass + graph capture), fills inputs and noise on the device
    torch.manual_seed(1)
    last_ticket = run_steps(max(3, args.warmup), resident=False)
    d2h_bytes = pipe.d2h_bytes(last_ticket)
    launches_per_step = pipe.launches_per_step()

    # ---- device-resident loop: inputs and noise already in HBM, graph replays only ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(pipe.slots[pipe.n % pipe.depth].stream)                     # the stream of the first timed step
    last_t = run_steps(args.steps, resident=True)
    e1.record(pipe.slots[last_t % pipe.depth].stream)                     # ... and of the last one
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    # two compute streams are in flight: bracket with device events and cross-check with the host clock
    ms_total = reduce_max(max(e0.elapsed_time(e1), wall_ms if args.steps >= 20 else 0.0))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end-to-end loop: every step copies its own pinned inputs H2D, draws the Sampler noise on the
    # CPU generator (reference semantics), and lands its own results in host memory, inside the timed region
    # (one compute stream shared by the two slots: with host copies in the loop, graphs that overlap only
    # partially slow each other down -- measured 11.8 vs 14.2 ms/step, scripts/e2e_probe.py)
    pipe_e2e = NovelViewPipeline(model, depth=3, shared_stream=True,
                                 gather_fn=(lambda c: gather_crops(c, world * B)) if world > 1 else None)
    run_steps(5, resident=False, pipe=pipe_e2e)
    e2e_steps = max(4, args.steps)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(pipe_e2e.copy_stream)
    run_steps(e2e_steps, resident=False, pipe=pipe_e2e)
    e1.record(pipe_e2e.out_stream)
    torch.cuda.synchronize()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = reduce_max(max(e0.elapsed_time(e1), e2e_wall))
    barrier()
    e2e_value = world * B / (e2e_ms / e2e_steps * 1e-3)
    del pipe_e2e

    # ---- the same loop with the VUNet inputs shipped as the three uint8 images the reference holds before to_tensor
    # (trajectory_inference.py:215-220; to_tensor / flip / concat run on the device): 9 instead of 36 bytes per pixel
    mk, nsrc, ndst = synth.make_vunet_inputs_u8(first, B)
    host_u8 = {k: v for k, v in host.items() if k not in ("x", "y")}
    host_u8.update(x_mask_u8=torch.from_numpy(mk).pin_memory(), x_normal_u8=torch.from_numpy(nsrc).pin_memory(),
                   y_normal_u8=torch.from_numpy(ndst).pin_memory())
    h2d_bytes_u8 = sum(v.numel() * v.element_size() for v in host_u8.values())
    pipe_u8 = NovelViewPipeline(model, depth=3, shared_stream=True,
                                gather_fn=(lambda c: gather_crops(c, world * B)) if world > 1 else None)
    run_steps(5, resident=False, pipe=pipe_u8, host=host_u8)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(pipe_u8.copy_stream)
    run_steps(e2e_steps, resident=False, pipe=pipe_u8, host=host_u8)
    e1.record(pipe_u8.out_stream)
    torch.cuda.synchronize()
    u8_wall = (time.perf_counter() - t0) * 1e3
    u8_ms = reduce_max(max(e0.elapsed_time(e1), u8_wall))
    barrier()
    e2e_u8_value = world * B / (u8_ms / e2e_steps * 1e-3)
    del pipe_u8

    # ---- per-launch roofline pass (CUDA events around every conv launch, on the launching stream)
    devin = {k: v.to(dev) for k, v in host.items()}
    noise_bank = {}

    def staged_noise(b, c, h, w):
        key = (b, c, h, w, staged_noise.i)
        staged_noise.i += 1
        if key not in noise_bank:
            noise_bank[key] = torch.randn((b, h, w, c), device=dev)
        return noise_bank[key]
    staged_noise.i = 0
    eng.noise_provider = staged_noise
    roof = None
    warp_roof = None
    if rank == 0:
        eng.profile = []
        staged_noise.i = 0
        model.fork_branches = False            # per-launch timing: one stream, no kernel overlaps another
        for _ in range(2):
            staged_noise.i = 0
            # park the stream behind a ~40 ms spin so the host enqueues the whole forward ahead of the GPU: the event
            # pairs then bracket kernel time, not host launch gaps (the narrow layers run 10-20 us, a launch costs ~15)
            torch.cuda._sleep(int(0.040 * 1.9e9))
            model(devin["y"], devin["x"])
        torch.cuda.synchronize()
        model.fork_branches = True
        recs = eng.profile
        eng.profile = None
        agg = {}
        for path, impl, flops, e0, e1 in recs:
            a = agg.setdefault(impl, [0.0, 0.0, 0])
            a[0] += flops
            a[1] += e0.elapsed_time(e1) * 1e-3
            a[2] += 1
        tc = agg.get(1, [0.0, 1e-9, 0])
        achieved = tc[0] / tc[1] / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_conv_traffic_summary.json")
        if os.path.exists(tpath) and B == 64:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["avg_traffic_bytes_per_launch"], tj["source"]
        roof = {"bound": "tensor", "kernel": "k_conv_tc (tcgen05 implicit-GEMM conv)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu; B=64)", "traffic_source": traffic_src, "peak_source": peaks["source"] + ", sustained bf16",
                "launches_per_step": tc[2] // 2, "avg_launch_ms": 1e3 * tc[1] / max(1, tc[2]),
                "algorithmic_flops_per_step": tc[0] / 2, "share_of_vunet_time": tc[1] / max(1e-9, sum(a[1] for a in agg.values()))}
        # fused warp kernel alone (BASELINE config 3 shape, HBM bound)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            warp_batch(devin["src"], devin["src_kp"], devin["dst_kp"], devin["K"], devin["E_src"], devin["E_dst"], devin["kp3d"], device=dev)
        reps = 20
        e0.record()
        for _ in range(reps):
            warp_batch(devin["src"], devin["src_kp"], devin["dst_kp"], devin["K"], devin["E_src"], devin["E_dst"], devin["kp3d"], device=dev)
        e1.record()
        torch.cuda.synchronize()
        wsec = e0.elapsed_time(e1) * 1e-3 / reps
        wgbs = B * WARP_BYTES_PER_CROP / wsec / 1e9
        warp_roof = {"bound": "hbm", "kernel": "fusg_warp_fused (k_visibility+k_homography+k_warp)", "achieved": wgbs, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": wgbs / peaks["hbm_gbs"], "crops": B, "ms": wsec * 1e3,
                     "note": f"{B} crops only ({B * WARP_BYTES_PER_CROP / 1e6:.0f} MB < L2, latency-bound); see large_batch"}
        # the same call at an HBM-sized batch: this rank's crops tiled to 4096 (4.8 GB of algorithmic traffic, the
        # compacted thread-per-solve homography path); BASELINE config 3 proper (16k crops) is scripts/bench_warp.py
        try:
            BL = 4096
            rep = (BL + B - 1) // B
            big = {k: devin[k].repeat((rep,) + (1,) * (devin[k].dim() - 1))[:BL].contiguous() for k in ("src", "src_kp", "dst_kp", "K", "E_src", "E_dst", "kp3d")}
            for _ in range(2):
                rb = warp_batch(big["src"], big["src_kp"], big["dst_kp"], big["K"], big["E_src"], big["E_dst"], big["kp3d"], device=dev)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                rb = warp_batch(big["src"], big["src_kp"], big["dst_kp"], big["K"], big["E_src"], big["E_dst"], big["kp3d"], device=dev)
            e1.record()
            torch.cuda.synchronize()
            bsec = e0.elapsed_time(e1) * 1e-3 / 3
            bgbs = BL * WARP_BYTES_PER_CROP / bsec / 1e9
            warp_roof["large_batch"] = {"crops": BL, "ms": bsec * 1e3, "achieved": bgbs, "unit": "GB/s", "frac": bgbs / peaks["hbm_gbs"]}
            del big, rb
            torch.cuda.empty_cache()
        except RuntimeError as exc:                       # e.g. not enough free memory next to the pipelines: report, do not fail the bench
            warp_roof["large_batch"] = {"skipped": str(exc)[:120]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(steps=8, warmup=1)
        cpu.pop("ms_per_step", None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{B} synthetic 256x256 vehicle crops per GPU per step: fused planar warp (5 CAD planes) + VUNet bf16 forward "
                                   f"(BASELINE config 2), random-init weights" + (", NCCL all-gather of completed uint8 crops" if world > 1 else ""),
                       "crops_per_gpu": B, "global_crops_per_step": world * B, "parallelism": f"crop-sharded dp{world}",
                       "l2": "per-step activations (>5 GB) exceed the 126 MB L2; no explicit flush",
                       "launch": "one CUDA-graph replay per step (gpu_launches counts the kernels inside the graphs)" + (" + eager NCCL all-gather" if world > 1 else "")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "note": "NovelViewPipeline.submit/result: pinned host inputs -> H2D, Sampler noise drawn on the CPU generator (reference "
                            "semantics), CUDA-graph replay, completed crops + warped planes + flags D2H; 3 slots (2 steps in flight behind the one being submitted), all inside the timed region"},
            "e2e_u8": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes_u8, "d2h_bytes_per_step": d2h_bytes,
                       "ms_per_step": u8_ms / e2e_steps, "steps": e2e_steps,
                       "note": "same loop, VUNet inputs shipped as the three uint8 images the reference holds before to_tensor "
                               "(trajectory_inference.py:215-220); to_tensor / channel flip / concat on the device (fusg_u8_to_vunet_inputs)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "roofline_warp": warp_roof,
            "vunet_tflops_whole_forward": world * B * VUNET_FLOPS_PER_CROP / (ms_per_step * 1e-3) / 1e12,
            "cpu_baseline": cpu,
        }
        if world == 1 and not args.no_icn:
            line["icn_generator"] = icn_info(torch, synth, B, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def icn_info(torch, synth, B, dev):
    """Informational, outside every timed region of the headline metric: the ICN generator G_Resnet (the consumer of
    the warped planes in the reference's data flow) on the same batch size -- crops/s and FLOP-weighted tensor throughput."""
    try:
        from future_urban_scene_generation_b200.warp_learn.models import G_Resnet
        torch.manual_seed(0)
        g = G_Resnet(21).to(dev).eval()
        xi = torch.from_numpy(synth.make_icn_inputs(0, min(B, 8), 256)).to(dev)
        xi = xi.repeat((B + xi.shape[0] - 1) // xi.shape[0], 1, 1, 1)[:B].contiguous()
        for _ in range(3):
            g(xi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g(xi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flops = 130124087296.0                       # per 256x256 crop, 18 convolutions (oracle/icn_oracle.py: flops_per_crop)
        return {"crops_per_s": B / ms * 1e3, "ms_per_forward": ms, "crops": B, "tflops_whole_forward": B * flops / (ms * 1e-3) / 1e12,
                "dtype": "fp16 operands, fp32 accumulate", "note": "not part of `value`; details in profiles/r1_icn_bench.json"}
    except Exception as ex:                          # never let the informational block take the bench line down
        return {"error": f"{type(ex).__name__}: {ex}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--crops-per-rank", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-icn", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that write to fd 1 on their own (NCCL prints its version there
    # whatever NCCL_DEBUG_FILE says) are pointed at stderr for the whole run, and print() gets the real stdout back
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
