set -x
cd /root/repo
mkdir -p gpurun_out/j12
micro() { name=$1; shift; env "$@" CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j12/micro_$name.txt 2>&1; python - <<PY
import json
print("$name", ' | '.join('%s %.0f'%(json.loads(l)['layer'].replace(' ',''), json.loads(l)['us']) for l in open('gpurun_out/j12/micro_$name.txt') if l.startswith('{')))
PY
}
(timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py -m gpu -q > gpurun_out/j12/pytest_conv.txt 2>&1; echo "rc=$?" >> gpurun_out/j12/pytest_conv.txt); tail -3 gpurun_out/j12/pytest_conv.txt
micro default IRN_F16_X=0
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j12/bench_c3.json 2> gpurun_out/j12/bench_c3.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/j12/bench_c3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_conv']['conv_path_ms_per_step'], d['roofline_conv']['frac'], d['roofline_conv']['frac_issued'], d['clocks'])"
which compute-sanitizer || export PATH=$PATH:/usr/local/cuda/bin
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10 --kernel-regex kns=conv_f16 python tools/sanitize_small.py > gpurun_out/j12/racecheck_f16.txt 2>&1; tail -4 gpurun_out/j12/racecheck_f16.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10 --kernel-regex kns=rw_ python tools/sanitize_small.py > gpurun_out/j12/racecheck_rw.txt 2>&1; tail -4 gpurun_out/j12/racecheck_rw.txt
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 --kernel-regex kns=conv_f16 python tools/sanitize_small.py > gpurun_out/j12/initcheck_f16.txt 2>&1; tail -4 gpurun_out/j12/initcheck_f16.txt
