# Round 2: mlp_mid.cu as the default for 9 .. 64 poses - the whole GPU suite, then the latency sweep.
mkdir -p gpurun_out
O=gpurun_out/r2mid2
timeout 100 python tools/bench_latency.py 9 16 17 32 33 48 64 65 > ${O}_lat.txt 2>&1; cat ${O}_lat.txt | tail -8
timeout 400 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -5 ${O}_tests.log
