set -x
cd /root/repo
mkdir -p gpurun_out/j25
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_steps.py -m gpu -q -k "spawn or batched" > gpurun_out/j25/pytest_spawn.txt 2>&1; echo "rc=$?" >> gpurun_out/j25/pytest_spawn.txt); tail -5 gpurun_out/j25/pytest_spawn.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/j25/bench_c3_n2.json 2> gpurun_out/j25/bench_c3_n2.err; echo rc=$?
IRN_STEP_PROFILE=1 timeout 600 $TR bench.py --gpus 2 --config 4 --steps 1 > gpurun_out/j25/bench_c4_n2.json 2> gpurun_out/j25/bench_c4_n2.err; echo rc=$?
grep "step profile" gpurun_out/j25/bench_c4_n2.err | tail -8 | sed 's/^.*\[irn/[irn/'
python - <<'PY'
import json
for f in ("c3_n2", "c4_n2"):
    d = json.load(open("gpurun_out/j25/bench_%s.json" % f))
    print(f, d["value"], d["ms_per_step"], d.get("per_rank_ms_per_step", d.get("per_rank_seconds")), d.get("rank0_pass_seconds"), d["clocks"])
PY
