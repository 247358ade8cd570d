set -x
cd /root/repo
mkdir -p gpurun_out/j15
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/j15/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/j15/pytest.txt)
tail -5 gpurun_out/j15/pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j15/smoke.txt 2>&1; tail -2 gpurun_out/j15/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j15/bench_c3.json 2> gpurun_out/j15/bench_c3.err; echo rc=$?
timeout 600 python bench.py --config 4 --steps 1 > gpurun_out/j15/bench_c4.json 2> gpurun_out/j15/bench_c4.err; echo rc=$?
cat gpurun_out/j15/bench_c3.json | cut -c1-400; python -c "
import json
d=json.load(open('gpurun_out/j15/bench_c4.json')); print(d['value'], d['rank0_pass_seconds'])"
