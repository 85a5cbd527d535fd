mkdir -p gpurun_out
T=r02u
timeout -k 5 120 python -m pytest tests/test_gpu_pair.py -x -q --timeout=40 -p no:cacheprovider > gpurun_out/${T}_pair.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pair.log
tail -5 gpurun_out/${T}_pair.log
if grep -q "rc=0" gpurun_out/${T}_pair.log; then
  for pr in 0 2; do FRIDO_TC_PAIR=$pr FRIDO_SK=0 timeout 120 python tools/prof/conv_bench.py 7 8 9 4 5 >> gpurun_out/${T}_convbench.log 2>&1; done
  cat gpurun_out/${T}_convbench.log
fi
