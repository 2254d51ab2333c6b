# Round 2: mlp_mid.cu - sentinel polling A/B with in-kernel stamps on a warm GPU, then parity on the default.
mkdir -p gpurun_out
O=gpurun_out/r2mid4
timeout 60 python tools/forward_once.py 1048576 40 > /dev/null 2>&1
for S in 1 0; do
  for B in 9 32 64; do
    P3D_MID_SENTINEL=$S P3D_LAT_STAMPS=1 timeout 60 python tools/bench_latency.py $B > ${O}_s${S}_$B.txt 2>&1; echo "sentinel=$S: $(tail -2 ${O}_s${S}_$B.txt | tr '\n' ' ')"
  done
done
timeout 60 python tools/bench_latency.py 9 16 17 32 33 48 64 65 > ${O}_lat.txt 2>&1; cat ${O}_lat.txt
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q -k "ragged or mid_batch" > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
