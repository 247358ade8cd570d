"""How fast is the CPU oracle's conv path on this host at different torch thread counts?  (Probe behind bench.py's choice of
32 threads for the cpu_baseline leg; lives under tests/ because it imports the oracle.  Run: python tests/probe_cpu_threads.py)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irn_b200 import synth
from oracle import nets
sd = synth.cam_state_dict()
x = synth.normalize_image(synth.image(0, 512, 512)); x = torch.from_numpy(np.stack([x, x[..., ::-1].copy()]))
for n in (8, 16, 32, 64, 128):
    if n > os.cpu_count(): break
    torch.set_num_threads(n)
    with torch.no_grad():
        nets.cam_forward(x, sd)
        t = time.perf_counter(); nets.cam_forward(x, sd); dt = time.perf_counter() - t
        a = torch.rand(4096, 4096); t = time.perf_counter(); torch.matmul(a, a); dm = time.perf_counter() - t
    print("threads", n, "cam512 %.2fs" % dt, "sgemm4096 %.2fs (%.2f TFLOP/s)" % (dm, 2 * 4096**3 / dm / 1e12), flush=True)
