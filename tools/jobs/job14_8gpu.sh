set -x
cd /root/repo
mkdir -p gpurun_out/j14
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR bench.py --gpus 8 --config 4 --steps 1 > gpurun_out/j14/bench_c4_n8.json 2> gpurun_out/j14/bench_c4_n8.err; echo rc=$?
timeout 600 $TR bench.py --gpus 8 --config 5 --steps 5 > gpurun_out/j14/bench_c5_n8.json 2> gpurun_out/j14/bench_c5_n8.err; echo rc=$?
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/j14/bench_c3_n8.json 2> gpurun_out/j14/bench_c3_n8.err; echo rc=$?
for f in c4 c5 c3; do tail -n 3 gpurun_out/j14/bench_${f}_n8.err | cut -c1-300; cat gpurun_out/j14/bench_${f}_n8.json | cut -c1-1500; done
