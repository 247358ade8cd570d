set -x
cd /root/repo
mkdir -p gpurun_out/j5
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_steps.py -m gpu -q -k "spawn or batched" > gpurun_out/j5/pytest_spawn.txt 2>&1; echo "rc=$?" >> gpurun_out/j5/pytest_spawn.txt); tail -5 gpurun_out/j5/pytest_spawn.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/j5/bench_c3_n2.json 2> gpurun_out/j5/bench_c3_n2.err; echo rc=$?
timeout 600 $TR bench.py --gpus 2 --config 4 --steps 1 > gpurun_out/j5/bench_c4_n2.json 2> gpurun_out/j5/bench_c4_n2.err; echo rc=$?
timeout 600 $TR bench.py --gpus 2 --config 5 --steps 3 > gpurun_out/j5/bench_c5_n2.json 2> gpurun_out/j5/bench_c5_n2.err; echo rc=$?
tail -3 gpurun_out/j5/*.err
cat gpurun_out/j5/*.json
