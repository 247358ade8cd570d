set -x
cd /root/repo
mkdir -p gpurun_out/j22
(timeout 900 python -m pytest tests/test_gpu_steps.py -m gpu -q > gpurun_out/j22/pytest_steps.txt 2>&1; echo "rc=$?" >> gpurun_out/j22/pytest_steps.txt); tail -4 gpurun_out/j22/pytest_steps.txt
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j22/c4.json 2> gpurun_out/j22/c4.err; echo rc=$?
grep "step profile" gpurun_out/j22/c4.err | tail -4 | sed 's/^.*\[irn/[irn/'
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j22/c5.json 2> gpurun_out/j22/c5.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/j22/c4.json')); print(d['value'], d['rank0_pass_seconds'], d['clocks'])
d=json.load(open('gpurun_out/j22/c5.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
