set -x
cd /root/repo
mkdir -p gpurun_out/j26
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/j26/bench_c3_n8.json 2> gpurun_out/j26/bench_c3_n8.err; echo rc=$?
timeout 400 $TR bench.py --gpus 8 --config 4 --steps 1 > gpurun_out/j26/bench_c4_n8.json 2> gpurun_out/j26/bench_c4_n8.err; echo rc=$?
python - <<'PY'
import json
for f in ("c3_n8", "c4_n8"):
    try:
        d = json.load(open("gpurun_out/j26/bench_%s.json" % f))
        print(f, d["value"], d["ms_per_step"], d.get("per_rank_ms_per_step", d.get("per_rank_seconds")), d.get("rank0_pass_seconds"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -n 3 gpurun_out/j26/*.err | cut -c1-300
