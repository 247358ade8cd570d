set -x
cd /root/repo
mkdir -p gpurun_out/j13
IRN_STEP_PROFILE=1 timeout 600 python bench.py --config 4 --steps 1 > gpurun_out/j13/bench_c4.json 2> gpurun_out/j13/bench_c4.err; echo rc=$?
grep "step profile" gpurun_out/j13/bench_c4.err | tail -2
python -c "
import json
d=json.load(open('gpurun_out/j13/bench_c4.json')); print(d['value'], d['rank0_pass_seconds'])"
which compute-sanitizer || export PATH=$PATH:/usr/local/cuda/bin
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 --kernel-regex kns=conv_f16 python tools/sanitize_small.py > gpurun_out/j13/initcheck_f16.txt 2>&1; tail -3 gpurun_out/j13/initcheck_f16.txt
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 --kernel-regex kns=conv_tc python tools/sanitize_small.py > gpurun_out/j13/initcheck_tc.txt 2>&1; tail -3 gpurun_out/j13/initcheck_tc.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 4 --kernel-regex kns=conv_tc_persist python tools/sanitize_small.py > gpurun_out/j13/racecheck_persist.txt 2>&1; tail -3 gpurun_out/j13/racecheck_persist.txt
(timeout 600 python -m pytest tests/test_gpu_steps.py -m gpu -q > gpurun_out/j13/pytest_steps.txt 2>&1; echo "rc=$?" >> gpurun_out/j13/pytest_steps.txt); tail -3 gpurun_out/j13/pytest_steps.txt
