set -x
cd /root/repo
mkdir -p gpurun_out/j24
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > gpurun_out/j24/clocks.csv &
SMI=$!
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j24/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/j24/pytest.txt)
tail -6 gpurun_out/j24/pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j24/smoke.txt 2>&1; tail -2 gpurun_out/j24/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j24/bench_c3.json 2> gpurun_out/j24/bench_c3.err; echo rc=$?
timeout 300 python bench.py --config 2 --steps 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/j24/bench_c2.json 2> gpurun_out/j24/bench_c2.err; echo rc=$?
timeout 300 python bench.py --config 5 --steps 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/j24/bench_c5.json 2> gpurun_out/j24/bench_c5.err; echo rc=$?
timeout 600 python bench.py --config 4 --steps 1 > gpurun_out/j24/bench_c4.json 2> gpurun_out/j24/bench_c4.err; echo rc=$?
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j24/bench_ref.json 2> gpurun_out/j24/bench_ref.err; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/j24/layer_metrics.csv python tools/net_forward_once.py 8 > gpurun_out/j24/layer_metrics.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/j24/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/j24/launches_bench.log 2>&1; echo rc=$?
kill $SMI
gzip -f gpurun_out/j24/layer_metrics.csv gpurun_out/j24/launches_bench.csv
for f in c3 c2 c5 c4 ref; do echo "== $f"; python - <<PY
import json
d = json.load(open("gpurun_out/j24/bench_$f.json"))
print({k: d[k] for k in d if k not in ("config",)})
PY
done
