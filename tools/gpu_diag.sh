#!/bin/bash
# Run the GPU checks in separate processes (a trapped kernel kills only that process' context),
# each under a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/summary.txt
timeout 300 python tools/diag_umma.py > gpurun_out/diag_umma.log 2>&1; echo "diag_umma rc=$?" | tee -a gpurun_out/summary.txt
tail -n 25 gpurun_out/diag_umma.log
rc_all=0
for f in tests/test_gpu_mlp.py tests/test_gpu_geometry.py tests/test_gpu_eval.py tests/test_gpu_train.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout=600 -x "$@" > gpurun_out/$name.log 2>&1
  rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 40 gpurun_out/$name.log
  [ $rc -ne 0 ] && rc_all=1
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/smoke.log
exit $rc_all
