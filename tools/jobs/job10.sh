set -x
cd /root/repo
mkdir -p gpurun_out/j10
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 > gpurun_out/j10/bench_c4.json 2> gpurun_out/j10/bench_c4.err; echo rc=$?
grep "step profile" gpurun_out/j10/bench_c4.err
python -c "
import json
d=json.load(open('gpurun_out/j10/bench_c4.json')); print(d['value'], d['rank0_pass_seconds'])"
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --num-workers 24 > gpurun_out/j10/bench_c4_w24.json 2> gpurun_out/j10/bench_c4_w24.err; echo rc=$?
grep "step profile" gpurun_out/j10/bench_c4_w24.err | tail -2
python -c "
import json
d=json.load(open('gpurun_out/j10/bench_c4_w24.json')); print(d['value'], d['rank0_pass_seconds'])"
nproc
