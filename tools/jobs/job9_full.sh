set -x
cd /root/repo
mkdir -p gpurun_out/j9
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > gpurun_out/j9/clocks.csv &
SMI=$!
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j9/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/j9/pytest.txt)
tail -8 gpurun_out/j9/pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j9/smoke.txt 2>&1; tail -2 gpurun_out/j9/smoke.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/j9/bench_c3.json 2> gpurun_out/j9/bench_c3.err; echo rc=$?
timeout 300 python bench.py --config 2 --steps 5 > gpurun_out/j9/bench_c2.json 2> gpurun_out/j9/bench_c2.err; echo rc=$?
timeout 300 python bench.py --config 5 --steps 5 > gpurun_out/j9/bench_c5.json 2> gpurun_out/j9/bench_c5.err; echo rc=$?
timeout 600 python bench.py --config 4 --steps 1 > gpurun_out/j9/bench_c4.json 2> gpurun_out/j9/bench_c4.err; echo rc=$?
timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/j9/bench_ref.json 2> gpurun_out/j9/bench_ref.err; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/j9/layer_metrics.csv python tools/net_forward_once.py 8 > gpurun_out/j9/layer_metrics.log 2>&1; echo rc=$?
CONV_MODE=2 CONV_ONLY="L3 c2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16_halo -s 6 -c 1 -o gpurun_out/j9/ncu_halo_L3c2 python tools/conv_micro.py > gpurun_out/j9/ncu.log 2>&1; echo rc=$?
CONV_MODE=2 CONV_ONLY="L4 c1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16_kernel -s 6 -c 1 -o gpurun_out/j9/ncu_f16_L4c1 python tools/conv_micro.py >> gpurun_out/j9/ncu.log 2>&1; echo rc=$?
CONV_MODE=2 CONV_ONLY="L1 c3" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16_kernel -s 6 -c 1 -o gpurun_out/j9/ncu_f16_L1c3 python tools/conv_micro.py >> gpurun_out/j9/ncu.log 2>&1; echo rc=$?
kill $SMI
for f in c3 c2 c5 c4 ref; do echo "== $f"; cat gpurun_out/j9/bench_$f.json | cut -c1-3000; tail -2 gpurun_out/j9/bench_$f.err; done
