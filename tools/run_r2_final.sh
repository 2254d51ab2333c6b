# Round 2, closing check of the shipped build (1 GPU): smoke + the inference parity module + the training golden cases.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2fin_smoke.log
timeout 200 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_realtime.py -x -q > gpurun_out/r2fin_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2fin_tests.log
timeout 60 python tools/bench_latency.py 1 8 9 16 32 64 65 256 > gpurun_out/r2fin_lat.txt 2>&1; cat gpurun_out/r2fin_lat.txt
