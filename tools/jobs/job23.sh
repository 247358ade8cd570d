set -x
cd /root/repo
mkdir -p gpurun_out/j23
(timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py -m gpu -q > gpurun_out/j23/pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/j23/pytest.txt); tail -4 gpurun_out/j23/pytest.txt
CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j23/micro_pf1.txt 2>&1
IRN_F16_RES_PREFETCH=0 CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j23/micro_pf0.txt 2>&1
python - <<'PY'
import json
def rd(f):
    return {json.loads(l)["layer"]: json.loads(l)["us"] for l in open(f) if l.startswith("{")}
a, b = rd("gpurun_out/j23/micro_pf1.txt"), rd("gpurun_out/j23/micro_pf0.txt")
for k in a: print("%-24s prefetch %7.1f  off %7.1f us" % (k, a[k], b.get(k, 0)))
PY
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j23/bench_pf1.json 2> gpurun_out/j23/bench_pf1.err; echo rc=$?
IRN_F16_RES_PREFETCH=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j23/bench_pf0.json 2> gpurun_out/j23/bench_pf0.err; echo rc=$?
python -c "
import json
for f in ('pf1','pf0'):
    d=json.load(open('gpurun_out/j23/bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline_conv']['conv_path_ms_per_step'], d['clocks'])"
