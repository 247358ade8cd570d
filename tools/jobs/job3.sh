set -x
cd /root/repo
mkdir -p gpurun_out/j3
run_variant() {   # name, env...
  name=$1; shift
  (env "$@" timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k f16 > gpurun_out/j3/pytest_$name.txt 2>&1; echo "rc=$?" >> gpurun_out/j3/pytest_$name.txt)
  tail -4 gpurun_out/j3/pytest_$name.txt
  env "$@" timeout 200 python tools/net_err.py 2 > gpurun_out/j3/neterr_$name.txt 2>&1
  cat gpurun_out/j3/neterr_$name.txt | tail -2
  env "$@" CONV_MODE=2 timeout 200 python tools/conv_micro.py > gpurun_out/j3/micro_$name.txt 2>&1
  cat gpurun_out/j3/micro_$name.txt
}
timeout 200 python tools/net_err.py 1 > gpurun_out/j3/neterr_mode1.txt 2>&1; cat gpurun_out/j3/neterr_mode1.txt | tail -1
run_variant halo_acc1 IRN_F16_HALO=1
run_variant halo_acc3 IRN_F16_HALO=1 IRN_F16_ACC3_MINK=512
run_variant nohalo_acc3 IRN_F16_HALO=0 IRN_F16_ACC3_MINK=512
run_variant halo_acc3_all IRN_F16_HALO=1 IRN_F16_ACC3_MINK=64
grep -h conv_f16x3 gpurun_out/parity_measured.jsonl | tail -60 > gpurun_out/j3/conv_records.txt
