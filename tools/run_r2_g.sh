mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_realtime.py tests/test_gpu_mlp.py -x -q > gpurun_out/r2g_tests.log 2>&1; tail -3 gpurun_out/r2g_tests.log
timeout 100 python tools/bench_realtime.py > gpurun_out/r2g_realtime.txt 2>&1; tail -3 gpurun_out/r2g_realtime.txt
P3D_LAT_GRIDLL=0 timeout 100 python tools/bench_realtime.py > gpurun_out/r2g_realtime_cluster.txt 2>&1; tail -2 gpurun_out/r2g_realtime_cluster.txt
