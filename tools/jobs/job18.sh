set -x
cd /root/repo
mkdir -p gpurun_out/j18
nproc
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j18/c4_w12.json 2> gpurun_out/j18/c4_w12.err; echo rc=$?
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --num-workers 24 --no-cpu-baseline --no-eager-baseline > gpurun_out/j18/c4_w24.json 2> gpurun_out/j18/c4_w24.err; echo rc=$?
grep "step profile" gpurun_out/j18/*.err | tail -20
python -c "
import json
for f in ('w12','w24'):
    d=json.load(open('gpurun_out/j18/c4_%s.json'%f)); print(f, d['value'], d['rank0_pass_seconds'], d['clocks'])"
