# (historical) A/B of gather variants of mlp_mid.cu; the switches used here (P3D_MID_BATCH / _ROTATE / _UNCOND) were removed once measured - results: profiles/r2_mid_batch_latency.txt
mkdir -p gpurun_out
O=gpurun_out/r2mid5
timeout 60 python tools/forward_once.py 1048576 40 > /dev/null 2>&1
for BT in 32 16 32 16; do
  P3D_MID_BATCH=$BT P3D_LAT_STAMPS=1 timeout 60 python tools/bench_latency.py 64 > ${O}_b${BT}.txt 2>&1; echo "batch=$BT: $(tail -2 ${O}_b${BT}.txt | tr '\n' ' ')"
done
