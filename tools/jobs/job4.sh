set -x
cd /root/repo
mkdir -p gpurun_out/j4
(timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k f16 > gpurun_out/j4/pytest_halo.txt 2>&1; echo "rc=$?" >> gpurun_out/j4/pytest_halo.txt); tail -3 gpurun_out/j4/pytest_halo.txt
(IRN_F16_HALO=0 IRN_F16_ACC3_MINK=512 timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k f16 > gpurun_out/j4/pytest_nohalo_acc3.txt 2>&1; echo "rc=$?" >> gpurun_out/j4/pytest_nohalo_acc3.txt); tail -3 gpurun_out/j4/pytest_nohalo_acc3.txt
for mk in 512 1024 2048 1000000; do
  IRN_F16_ACC3_MINK=$mk timeout 200 python tools/net_err.py 2 2>&1 | tail -1 > gpurun_out/j4/neterr_acc3_$mk.txt; cat gpurun_out/j4/neterr_acc3_$mk.txt
  IRN_F16_ACC3_MINK=$mk CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j4/micro_acc3_$mk.txt 2>&1; cat gpurun_out/j4/micro_acc3_$mk.txt
done
IRN_F16_HALO=0 CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j4/micro_nohalo_acc1.txt 2>&1; cat gpurun_out/j4/micro_nohalo_acc1.txt
IRN_F16_ACC3_MINK=512 timeout 400 python bench.py --steps 5 --warmup 3 --conv-mode 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/j4/bench_c3_acc3_512.json 2> gpurun_out/j4/bench_c3_acc3_512.err; echo rc=$?
IRN_F16_ACC3_MINK=2048 timeout 400 python bench.py --steps 5 --warmup 3 --conv-mode 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/j4/bench_c3_acc3_2048.json 2> gpurun_out/j4/bench_c3_acc3_2048.err; echo rc=$?
IRN_F16_ACC3_MINK=512 CONV_MODE=2 CONV_ONLY="L3 c2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16_halo -s 6 -c 1 -o gpurun_out/j4/ncu_halo_L3c2 python tools/conv_micro.py > gpurun_out/j4/ncu.log 2>&1; echo rc=$?
IRN_F16_ACC3_MINK=512 CONV_MODE=2 CONV_ONLY="L1 c2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16_halo -s 6 -c 1 -o gpurun_out/j4/ncu_halo_L1c2 python tools/conv_micro.py >> gpurun_out/j4/ncu.log 2>&1; echo rc=$?
