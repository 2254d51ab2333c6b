# (historical) A/B of gather variants of mlp_mid.cu; the switches used here (P3D_MID_BATCH / _ROTATE / _UNCOND) were removed once measured - results: profiles/r2_mid_batch_latency.txt
# Round 2: mlp_mid.cu - warm GPU first, then A/B of the poll batch (16 / 32 poses) with in-kernel stamps; launch durations from ncu last.
mkdir -p gpurun_out
O=gpurun_out/r2mid3
timeout 60 python tools/forward_once.py 1048576 40 > /dev/null 2>&1
for BT in 16 32; do
  for B in 16 32 64; do
    P3D_MID_BATCH=$BT P3D_LAT_STAMPS=1 timeout 60 python tools/bench_latency.py $B > ${O}_b${BT}_$B.txt 2>&1; echo "batch=$BT: $(tail -2 ${O}_b${BT}_$B.txt | tr '\n' ' ')"
  done
done
P3D_MID_GRID=0 timeout 60 python tools/bench_latency.py 64 > ${O}_off_64.txt 2>&1; tail -1 ${O}_off_64.txt
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mid_grid -s 100 -c 6 --csv --log-file ${O}_ncu.csv python tools/bench_latency.py 64 > ${O}_ncu.log 2>&1; grep -c mid_grid ${O}_ncu.csv; tail -3 ${O}_ncu.csv | cut -c1-60,200-
