#!/usr/bin/env python
"""bench.py -- pseudo-label images/s of the IRN hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic VOC-shaped images per GPU:
multi-scale ResNet-50 CAM forward (scales 1.0/0.5/1.5/2.0, image + flip) -> CAM merge -> IRNet
EdgeDisplacement -> 256-iteration affinity random walk -> x4 upsample / argmax label map
(BASELINE.json configs[2], which contains configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); images shard across ranks with no data-path
collective, NCCL only gathers the per-image label maps at the end of each step (weak scaling).
Rank 0 prints ONE JSON line.  `--impl reference` times the CPU oracle port of the reference's own
algorithm (dense (hw)^2 transition matrix squared 8 times) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pseudo-label images/sec (CAM+random-walk)"
UNIT = "images/s"
H = W = 512
SCALES = (1.0, 0.5, 1.5, 2.0)
GFLOP_PER_IMAGE = 974.04 + 149.61       # 4-scale CAM + EdgeDisplacement


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-baseline-images", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--conv-mode", type=int, default=1, help="1 = tcgen05 3xTF32 (default), 0 = SIMT fp32")
    return ap.parse_args()


def config(batch, n_gpus):
    return {"workload": "batch=%d synthetic 512x512 per GPU: multi-scale CAM (0.5/1.0/1.5/2.0, image+flip) -> IRNet edge -> "
                        "256-iter random walk -> sem-seg label (BASELINE.json configs[2])" % batch,
            "global_batch": batch * n_gpus, "image": [H, W], "scales": list(SCALES), "rw_iters": 256, "beta": 10,
            "parallelism": "dp%d (images sharded, NCCL gather of label maps)" % n_gpus,
            "inputs": "decoded uint8 images [batch,512,512,3]; the 4-scale bicubic / normalise / flip pyramids (C1) are built on the device "
                      "inside the timed region",
            "l2": "every step writes and re-reads %.1f GB of fp32 pyramids plus the activations between two reads of the inputs: "
                  "far beyond the 126 MB L2, nothing survives from one step to the next" % (batch * 47.2e6 / 1e9),
            "weights": "seeded synthetic checkpoints in the reference's state_dict format (irn_b200/synth.py)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_threads():
    # measured on the B200 host (profiles/r01_cpu_threads_probe.txt): 128 torch threads are 50x slower than 16-32
    return min(os.cpu_count(), int(os.environ.get("IRN_CPU_THREADS", "32")))


def dense_walk_sample(cam, edge):
    """Time ONE of the reference's 8 dense squarings (misc/indexing.py:136-137) on the real 16384^2 transition matrix of
    this image plus the densify / normalise / final product, i.e. the walk with exp_times=1, and scale the squaring by 8."""
    import torch
    from oracle import indexing as oi
    t0 = time.perf_counter()
    oi.propagate_to_edge(cam, edge, 5, 10, 0)
    t_setup = time.perf_counter() - t0          # PathIndex + affinity + densify + normalise + x@T
    t0 = time.perf_counter()
    oi.propagate_to_edge(cam, edge, 5, 10, 1)
    t_one = time.perf_counter() - t0 - t_setup  # one squaring
    return t_setup, max(t_one, 1e-3)


def cpu_baseline_and_parity(out, ids, labels):
    """One image of the batch through the oracle port on the host cores.  CAM (4 scales), EdgeDisplacement and the label
    tail run in full; the dense walk is sampled (setup + 1 of 8 equal-cost squarings, x8) so the leg stays bounded.  The
    label map for the parity check uses the oracle's float64 stencil walk (the exact operator)."""
    import torch
    from irn_b200 import synth
    from oracle import pipeline as opipe
    from oracle import steps as osteps
    torch.set_num_threads(cpu_threads())
    torch.set_flush_denormal(True)
    cam_sd, irn_sd = synth.cam_state_dict(), synth.irn_state_dict()
    lab, t, aux = opipe.pseudo_label(synth.image(ids[0], H, W), labels[0], cam_sd, irn_sd, SCALES, walk="stencil")
    t_setup, t_sq = dense_walk_sample(aux["cam"], aux["edge"])
    total = t["preprocess"] + t["cam"] + t["irn"] + t_setup + 8 * t_sq + t["labels"]
    got = out["labels"][0].cpu().numpy()
    _, miou = osteps.confusion_miou([got], [lab])
    cpu = {"value": 1.0 / total, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
           "sample": "1 image of the batch: preprocess %.2fs + 4-scale CAM %.2fs + EdgeDisplacement %.2fs + dense walk (setup %.2fs + "
                     "8 x one measured 16384^2 fp32 squaring %.2fs) + labels %.2fs; torch CPU, flush-denormal on" %
                     (t["preprocess"], t["cam"], t["irn"], t_setup, t_sq, t["labels"])}
    parity = {"label_agreement_vs_oracle": float((lab == got).mean()), "miou_vs_oracle_labels": miou, "images": 1,
              "oracle_walk": "float64 stencil (exact operator)"}
    return cpu, parity


def run_reference(a, rank, out_stream):
    """CPU oracle port of the reference path, one image per step (bounded sample), all host threads."""
    import torch
    if rank != 0:
        return
    from irn_b200 import synth
    from oracle import pipeline as opipe
    torch.set_num_threads(cpu_threads())
    torch.set_flush_denormal(True)     # the favourable setting for the reference's dense squarings
    cam_sd, irn_sd = synth.cam_state_dict(), synth.irn_state_dict()
    steps, warm = min(a.steps, 2), 0   # one image is minutes of CPU work: bounded so the run ends in minutes
    tot = 0.0
    parts = {}
    for i in range(steps):
        _, t, aux = opipe.pseudo_label(synth.image(i, H, W), synth.label(i), cam_sd, irn_sd, SCALES, walk="stencil")
        t_setup, t_sq = dense_walk_sample(aux["cam"], aux["edge"])
        t["walk"] = t_setup + 8 * t_sq          # the reference's dense walk: setup + 8 equal-cost squarings, one measured
        tot += sum(t.values())
        for k, v in t.items():
            parts[k] = parts.get(k, 0.0) + v / steps
    dt = tot
    val = steps / dt
    sample = "1 image/step: PIL 4-scale preprocessing + 4-scale CAM (torch CPU fp32) + EdgeDisplacement + dense 256-step walk " \
             "(setup + 8 x one measured fp32 squaring of the 16384^2 transition matrix, flush-denormal on) + labels; stage seconds %s" % \
             {k: round(v, 3) for k, v in parts.items()}
    cfg = config(a.batch, a.gpus)      # the workload of the CUDA arm; every step here walks a bounded sample of it
    cfg["inputs"] = "decoded uint8 images; pyramids built by PIL on the host, as the reference's loader does"
    cfg["l2"] = "n/a (host run)"
    cfg["sample"] = "1 image of the batch per step"
    out_stream.write(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
                      "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": cfg,
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}) + "\n")
    out_stream.flush()


def clocks_mhz(clocks):
    try:
        return float(clocks.get("sm_mhz") or 1800.0)
    except Exception:
        return 1800.0


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner there when NCCL_DEBUG is set)
    write to file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to
    a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    a = parse()
    out_stream = _claim_stdout()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank, out_stream)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if not os.path.exists(entry.LIB):
        entry.build()
    from irn_b200 import _lib, synth
    from irn_b200.cam import CAM
    from irn_b200.irn import EdgeDisplacement
    from irn_b200.pipeline import PseudoLabelPipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    cam = CAM()
    cam.load_state_dict(synth.cam_state_dict(), strict=True)
    irn = EdgeDisplacement()
    irn.load_state_dict(synth.irn_state_dict(), strict=False)
    cam.cuda(dev), irn.cuda(dev)
    for m in (cam, irn):
        _lib.check(L.irn_net_set_conv_mode(m._get_plan(dev).handle, a.conv_mode))
    pipe = PseudoLabelPipeline(cam, irn, dev, SCALES, rw_sub_batch=64)

    # ---- synthetic inputs: rank r takes images r, r+N, ... of the global list (misc/torchutils.py:66-68)
    B = a.batch
    ids = [rank + world * i for i in range(B)]
    labels = np.stack([synth.label(i) for i in ids])
    host_inputs = torch.from_numpy(np.stack([synth.image(i, H, W) for i in ids])).pin_memory()     # uint8 [B,H,W,3]
    dev_inputs = host_inputs.to(dev)
    h2d_bytes = int(host_inputs.numel())
    d2h_bytes = B * H * W
    host_labels = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    gathered = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(world)] if world > 1 else None

    def step(from_host):
        out = pipe.run_u8(host_inputs if from_host else dev_inputs, labels)
        if world > 1:
            dist.all_gather(gathered, out["labels"])      # the one collective: per-image outputs to every rank
        if from_host:
            host_labels.copy_(out["labels"], non_blocking=True)
        return out

    def timed(n_steps, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(n_steps):
            out = step(from_host)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 0.0)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out

    for _ in range(max(a.warmup, 3)):
        step(False)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    L.irn_rw_set_timing(1)
    launches0 = L.irn_total_launch_count()
    ms_dev, wall_dev, out = timed(a.steps, False)
    launches = int(L.irn_total_launch_count() - launches0)
    import ctypes
    step_ms, n_it = ctypes.c_float(), ctypes.c_int()
    have_rw = L.irn_rw_last_step_ms(ctypes.byref(step_ms), ctypes.byref(n_it)) == 0
    L.irn_rw_set_timing(0)
    step(True)                                   # warm the host path (pinned copies)
    ms_e2e, wall_e2e, out = timed(a.steps, True)
    clocks = sampler.stop()

    value = world * B * a.steps / (ms_dev / 1e3)
    e2e = world * B * a.steps / (max(ms_e2e, wall_e2e) / 1e3)   # host-side time counts for the end-to-end number

    # ---- roofline of the walk's step kernel (last walk of the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    if have_rw:
        counts = [len(k) for k in out["keys"]]
        last = counts[-(len(counts) % pipe.rw_sub or pipe.rw_sub):]
        n_img, totc, N = len(last), sum(last), (H // 4) * (W // 4)
        alg_step = N * (n_img * (4 * 34 + 8) + 2 * 8 * totc)      # per walk step: fp32 weights + fp64 1/s + fp64 state read+write
        src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"
        if L.irn_rw_last_was_fused():
            # the whole walk is ONE launch: algorithmic bytes = n_iter x the per-step figure (SURVEY.md 8(d) B_rw)
            launch_ms = step_ms.value * n_it.value
            alg = alg_step * n_it.value
            traffic = None
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", "rw_fused_traffic.json")))
                traffic = int(t["dram_bytes_per_item"] * totc)
            except Exception:
                pass
            ach = alg / (launch_ms * 1e-3) / 1e9
            # the bound the resident design actually runs into: shared-memory wavefronts
            wf_per_warp_step = (34 * 2 - 16) * 4 + 108 * 2      # weight LDS.32 (16 planes' forward taps come from registers) + state LDS.64 x 2 wavefronts
            clusters = max(1, int(L.irn_rw_last_was_fused()))
            per = -(-totc // clusters)                         # items walked by the busiest cluster
            smem_cycles = wf_per_warp_step * 8 * per * n_it.value   # 8 warps per CTA, one 128-byte wavefront per cycle per SM
            roofline = {"kernel": "rw_fused_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                        "traffic": traffic, "peak_source": src, "launch_us": 1e3 * launch_ms, "steps_per_launch": n_it.value,
                        "images_per_launch": n_img, "channels_per_launch": totc, "algorithmic_bytes_per_launch": alg,
                        "note": "weights stay resident in shared memory for all steps of a launch, so DRAM traffic (`traffic`) is a small "
                                "fraction of the algorithmic bytes and frac may exceed 1; the kernel's real ceiling is shared-memory bandwidth",
                        "clusters": clusters, "smem_wavefront_frac": smem_cycles / (launch_ms * 1e-3 * clocks_mhz(clocks) * 1e6)}
        else:
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "rw_step_traffic.json")))["dram_bytes_per_launch"]
            except Exception:
                pass
            ach = alg_step / (step_ms.value * 1e-3) / 1e9
            roofline = {"kernel": "rw_step_tma_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                        "traffic": traffic, "peak_source": src, "launch_us": 1e3 * step_ms.value, "images_per_launch": n_img,
                        "channels_per_launch": totc, "algorithmic_bytes_per_launch": alg_step}
    tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
    ach_tf = world * B * a.steps * GFLOP_PER_IMAGE / (ms_dev / 1e3) / 1e3 / world
    roofline_conv = {"bound": "tensor", "achieved": ach_tf, "unit": "TFLOP/s", "peak": tf32_peak, "frac": ach_tf / tf32_peak,
                     "note": "algorithmic conv FLOPs (1123.65 GFLOP/image) / whole-step time per GPU; 3xTF32 issues 3x these; "
                             "peak = measured sustained bf16 cuBLAS / 2 (tf32 runs at half the bf16 rate)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_dev / a.steps, "wall_ms_per_step": wall_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config(B, world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": max(ms_e2e, wall_e2e) / a.steps, "event_ms_per_step": ms_e2e / a.steps, "wall_ms_per_step": wall_e2e / a.steps,
                    "api": "PseudoLabelPipeline.run_u8 on pinned host uint8 images (decoded JPEGs); label maps copied back to pinned host memory"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_conv": roofline_conv,
            "conv_mode": "tcgen05 3xTF32" if a.conv_mode == 1 else "SIMT fp32"}

    # ---- CPU baseline + parity on a bounded sample (rank 0, single-GPU runs only)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"], line["parity"] = cpu_baseline_and_parity(out, ids, labels)
    if rank == 0:
        out_stream.write(json.dumps(line) + "\n")
        out_stream.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
