set -x
cd /root/repo
mkdir -p gpurun_out/j21
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j21/c4_w12.json 2> gpurun_out/j21/c4_w12.err; echo rc=$?
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 --warmup 3 --num-workers 6 --no-cpu-baseline --no-eager-baseline > gpurun_out/j21/c4_w6.json 2> gpurun_out/j21/c4_w6.err; echo rc=$?
for f in w12 w6; do grep "step profile" gpurun_out/j21/c4_$f.err | tail -4 | sed 's/^.*\[irn/[irn/'; done
python -c "
import json
for f in ('w12','w6'):
    d=json.load(open('gpurun_out/j21/c4_%s.json'%f)); print(f, d['value'], d['rank0_pass_seconds'], d['clocks'])"
