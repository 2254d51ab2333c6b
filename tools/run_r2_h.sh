# Round 2, GPU call H (1 GPU): persisting-L2 A/B of the fused inference kernel (time + DRAM traffic) and of the width-4096
# stress variant at 2^18 and 2^21 poses.
mkdir -p gpurun_out
O=gpurun_out/r2h
for P in 1 0; do
  P3D_L2_PERSIST=$P timeout 100 python tools/forward_once.py 1048576 20 > ${O}_fwd_persist$P.txt 2>&1; tail -1 ${O}_fwd_persist$P.txt
done
P3D_L2_PERSIST=0 timeout 150 python tools/capture_traffic.py > ${O}_traffic_persist0.txt 2>&1; tail -1 ${O}_traffic_persist0.txt; cp gpurun_out/mlp_tc_traffic.json ${O}_traffic_persist0.json
for P in 1 0; do
  for LG in 18 21; do
    P3D_L2_PERSIST=$P timeout 100 python tools/stress_once.py $LG 3 > ${O}_stress_p${P}_$LG.txt 2>&1; tail -1 ${O}_stress_p${P}_$LG.txt
  done
done
