set -x
cd /root/repo
mkdir -p gpurun_out/j1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > gpurun_out/j1/clocks.csv &
SMI=$!
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j1/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/j1/pytest.txt)
tail -15 gpurun_out/j1/pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/j1/bench_c3.json 2> gpurun_out/j1/bench_c3.err; echo rc=$?
timeout 300 python bench.py --config 2 --steps 3 > gpurun_out/j1/bench_c2.json 2> gpurun_out/j1/bench_c2.err; echo rc=$?
timeout 300 python bench.py --config 5 --steps 3 > gpurun_out/j1/bench_c5.json 2> gpurun_out/j1/bench_c5.err; echo rc=$?
timeout 600 python bench.py --config 4 --steps 1 --list-limit 1323 > gpurun_out/j1/bench_c4.json 2> gpurun_out/j1/bench_c4.err; echo rc=$?
# per-launch metrics over one forward (second pass): skip the launches of pass 0
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/j1/layer_metrics.csv python tools/net_forward_once.py 8 > gpurun_out/j1/layer_metrics.log 2>&1; echo rc=$?
# full captures of the long-K kernel: L4 3x3 and 2048->512
CONV_ONLY="L4 c2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ts_kernel -s 6 -c 1 -o gpurun_out/j1/ncu_ts_L4c2 python tools/conv_micro.py > gpurun_out/j1/ncu_ts.log 2>&1; echo rc=$?
CONV_ONLY="L4 c1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ts_kernel -s 6 -c 1 -o gpurun_out/j1/ncu_ts_L4c1 python tools/conv_micro.py >> gpurun_out/j1/ncu_ts.log 2>&1; echo rc=$?
kill $SMI
ls -la gpurun_out/j1
