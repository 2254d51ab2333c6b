# (historical) A/B of gather variants of mlp_mid.cu; the switches used here (P3D_MID_BATCH / _ROTATE / _UNCOND) were removed once measured - results: profiles/r2_mid_batch_latency.txt
mkdir -p gpurun_out
O=gpurun_out/r2mid8
timeout 60 python tools/forward_once.py 1048576 40 > /dev/null 2>&1
for C in "0 1" "0 0" "1 0" "0 1" "1 0"; do
  set -- $C
  P3D_MID_SENTINEL=$1 P3D_MID_UNCOND=$2 timeout 60 python tools/bench_latency.py 16 32 48 64 > ${O}_s$1_u$2.txt 2>&1; echo "sentinel=$1 uncond=$2: $(grep -o 'B=[0-9]*: p50 device [0-9.]* us\|back-to-back [0-9.]*' ${O}_s$1_u$2.txt | tr '\n' ' ')"
done
