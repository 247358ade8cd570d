set -x
cd /root/repo
mkdir -p gpurun_out/j2
timeout 120 python tools/bf16_bringup.py > gpurun_out/j2/bringup_swap0.txt 2>&1; echo rc=$?
IRN_BF_SWAP=1 timeout 120 python tools/bf16_bringup.py > gpurun_out/j2/bringup_swap1.txt 2>&1; echo rc=$?

cat gpurun_out/j2/bringup_swap0.txt gpurun_out/j2/bringup_swap1.txt gpurun_out/j2/bringup_nacc3.txt
(timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py tests/test_gpu_steps512.py tests/test_gpu_steps.py -m gpu -q > gpurun_out/j2/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/j2/pytest.txt)
tail -25 gpurun_out/j2/pytest.txt
CONV_MODE=2 timeout 300 python tools/conv_micro.py > gpurun_out/j2/conv_micro_f16.txt 2>&1; echo rc=$?
CONV_MODE=1 timeout 300 python tools/conv_micro.py > gpurun_out/j2/conv_micro_tf32.txt 2>&1; echo rc=$?
timeout 400 python bench.py --steps 5 --warmup 3 --conv-mode 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/j2/bench_c3_f16.json 2> gpurun_out/j2/bench_c3_f16.err; echo rc=$?
cat gpurun_out/j2/conv_micro_f16.txt gpurun_out/j2/conv_micro_tf32.txt
