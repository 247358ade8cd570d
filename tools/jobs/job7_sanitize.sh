set -x
cd /root/repo
mkdir -p gpurun_out/j7
which compute-sanitizer || export PATH=$PATH:/usr/local/cuda/bin
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/j7/memcheck.txt 2>&1; echo "rc=$?" >> gpurun_out/j7/memcheck.txt
tail -15 gpurun_out/j7/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/sanitize_small.py > gpurun_out/j7/racecheck.txt 2>&1; echo "rc=$?" >> gpurun_out/j7/racecheck.txt
tail -15 gpurun_out/j7/racecheck.txt
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/j7/initcheck.txt 2>&1; echo "rc=$?" >> gpurun_out/j7/initcheck.txt
tail -8 gpurun_out/j7/initcheck.txt
