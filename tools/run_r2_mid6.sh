# (historical) A/B of gather variants of mlp_mid.cu; the switches used here (P3D_MID_BATCH / _ROTATE / _UNCOND) were removed once measured - results: profiles/r2_mid_batch_latency.txt
mkdir -p gpurun_out
O=gpurun_out/r2mid6
timeout 60 python tools/forward_once.py 1048576 40 > /dev/null 2>&1
for R in 1 0 1 0; do
  for B in 32 64; do
    P3D_MID_ROTATE=$R P3D_LAT_STAMPS=1 timeout 60 python tools/bench_latency.py $B > ${O}_r${R}_$B.txt 2>&1; echo "rotate=$R: $(tail -2 ${O}_r${R}_$B.txt | tr '\n' ' ')"
  done
done
