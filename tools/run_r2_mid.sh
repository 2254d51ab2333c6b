# Round 2: the 9 .. 64 pose whole-chip kernel (mlp_mid.cu) - parity on that path, then latency A/B against the per-layer GEMMs.
mkdir -p gpurun_out
O=gpurun_out/r2mid
P3D_MID_GRID=1 timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q -k "ragged or mid_batch" > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -15 ${O}_tests.log
P3D_MID_GRID=1 timeout 100 python tools/bench_latency.py 9 16 17 32 33 64 > ${O}_lat_on.txt 2>&1; cat ${O}_lat_on.txt | tail -7
P3D_MID_GRID=0 timeout 100 python tools/bench_latency.py 9 16 17 32 33 64 > ${O}_lat_off.txt 2>&1; cat ${O}_lat_off.txt | tail -7
