set -x
cd /root/repo
mkdir -p gpurun_out/j8
(timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/j8/pytest_conv.txt 2>&1; echo "rc=$?" >> gpurun_out/j8/pytest_conv.txt); tail -3 gpurun_out/j8/pytest_conv.txt
timeout 200 python tools/net_err.py 2 2>&1 | tail -1 > gpurun_out/j8/neterr_default.txt; cat gpurun_out/j8/neterr_default.txt
micro() { name=$1; shift; env "$@" CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j8/micro_$name.txt 2>&1; python - <<PY
import json
print("$name", ' | '.join('%s %.0f'%(json.loads(l)['layer'].replace(' ',''), json.loads(l)['us']) for l in open('gpurun_out/j8/micro_$name.txt') if l.startswith('{')))
PY
}
micro default IRN_F16_X=0
micro nospin IRN_F16_SPIN=0
micro nohalo IRN_F16_HALO=0
micro acc_none IRN_F16_ACC_MINK=100000000
micro acc_all IRN_F16_ACC_MINK=64
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/j8/bench_c3.json 2> gpurun_out/j8/bench_c3.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/j8/bench_c3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_conv']['conv_path_ms_per_step'], d['roofline_conv']['frac'], d['roofline_conv']['frac_issued'], d['clocks'])"
(timeout 600 python -m pytest tests/test_gpu_jpeg.py tests/test_gpu_nets.py tests/test_gpu_steps512.py -m gpu -q > gpurun_out/j8/pytest_more.txt 2>&1; echo "rc=$?" >> gpurun_out/j8/pytest_more.txt); tail -15 gpurun_out/j8/pytest_more.txt
