set -x
cd /root/repo
mkdir -p gpurun_out/j17
(timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_steps512.py tests/test_gpu_conv.py -m gpu -q > gpurun_out/j17/pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/j17/pytest.txt); tail -8 gpurun_out/j17/pytest.txt
timeout 200 python tools/net_err.py 2 2>&1 | tail -1 | cut -c1-700
IRN_F16_STEM=0 timeout 200 python tools/net_err.py 2 2>&1 | tail -1 | cut -c1-700
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j17/bench_stem1.json 2> gpurun_out/j17/bench_stem1.err; echo rc=$?
IRN_F16_STEM=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j17/bench_stem0.json 2> gpurun_out/j17/bench_stem0.err; echo rc=$?
python -c "
import json
for f in ('stem1','stem0'):
    d=json.load(open('gpurun_out/j17/bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline_conv']['conv_path_ms_per_step'], d['gpu_launches'], d['clocks'])"
