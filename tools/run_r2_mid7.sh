# (historical) A/B of gather variants of mlp_mid.cu; the switches used here (P3D_MID_BATCH / _ROTATE / _UNCOND) were removed once measured - results: profiles/r2_mid_batch_latency.txt
mkdir -p gpurun_out
O=gpurun_out/r2mid7
timeout 60 python tools/forward_once.py 1048576 40 > /dev/null 2>&1
for U in 1 0 1 0; do
  P3D_MID_UNCOND=$U timeout 60 python tools/bench_latency.py 17 32 64 > ${O}_u${U}.txt 2>&1; echo "uncond=$U short warm-up: $(grep -o 'B=[0-9]*: p50 device [0-9.]* us\|back-to-back [0-9.]*' ${O}_u${U}.txt | tr '\n' ' ')"
done
timeout 100 python tools/forward_once.py 1048576 3000 > /dev/null 2>&1
for U in 1 0 1 0; do
  P3D_MID_UNCOND=$U timeout 60 python tools/bench_latency.py 17 32 64 > ${O}_u${U}w.txt 2>&1; echo "uncond=$U after 20 s of load: $(grep -o 'B=[0-9]*: p50 device [0-9.]* us\|back-to-back [0-9.]*' ${O}_u${U}w.txt | tr '\n' ' ')"
done
P3D_MID_GRID=0 timeout 60 python tools/bench_latency.py 64 | tail -1
