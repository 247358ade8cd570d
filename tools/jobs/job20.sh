set -x
cd /root/repo
mkdir -p gpurun_out/j20
(timeout 900 python -m pytest tests/test_gpu_steps.py tests/test_gpu_jpeg.py -m gpu -q > gpurun_out/j20/pytest_steps.txt 2>&1; echo "rc=$?" >> gpurun_out/j20/pytest_steps.txt); tail -8 gpurun_out/j20/pytest_steps.txt
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j20/c4.json 2> gpurun_out/j20/c4.err; echo rc=$?
grep "step profile" gpurun_out/j20/*.err | tail -2
python -c "
import json
d=json.load(open('gpurun_out/j20/c4.json')); print(d['value'], d['rank0_pass_seconds'], d['clocks'])"
