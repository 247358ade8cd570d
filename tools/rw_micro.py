"""Micro-benchmark of the random-walk path on one GPU (development aid; bench.py is the contract)."""
import argparse
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from irn_b200 import indexing, synth, _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-img", type=int, default=64)
    ap.add_argument("--c", type=int, default=2)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--iters", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--mix", action="store_true", help="per-image channel counts of bench.py's synthetic batch (1..3 classes)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    h = w = a.hw
    edges = torch.from_numpy(np.concatenate([synth.edge_map(h, w, "bimodal", i) for i in range(a.n_img)], 0)).to(dev)
    counts = [int(synth.label(i).sum()) for i in range(a.n_img)] if a.mix else [a.c] * a.n_img
    offs = np.concatenate([[0], np.cumsum(counts)])
    x = torch.rand((int(offs[-1]), h, w), device=dev)
    for _ in range(2):
        indexing.random_walk_batch(x, edges, offs, n_iter=a.iters, variant=a.variant)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        indexing.random_walk_batch(x, edges, offs, n_iter=a.iters, variant=a.variant)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    N = h * w
    e = 8
    totc = int(offs[-1])
    alg = a.iters * N * (a.n_img * (4 * 34 + 8) + 2 * e * totc)      # step launches only (bench.py's definition)
    print(json.dumps({"n_img": a.n_img, "C": a.c if not a.mix else "mix(%d)" % totc, "hw": a.hw, "iters": a.iters, "variant": a.variant, "ms": ms,
                      "ms_per_image": ms / a.n_img, "us_per_iter": 1e3 * ms / max(a.iters, 1),
                      "alg_GBps": alg / ms / 1e6, "launches": _lib.lib().irn_rw_last_launch_count()}))


if __name__ == "__main__":
    main()
