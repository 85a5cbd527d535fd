mkdir -p gpurun_out
T=r02F
timeout -k 5 200 python -m pytest tests/test_gpu_flash.py -x -q --timeout=40 -p no:cacheprovider > gpurun_out/${T}_flash.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_flash.log
tail -12 gpurun_out/${T}_flash.log
if grep -q "rc=0" gpurun_out/${T}_flash.log; then
  for pr in 0 1; do FRIDO_FLASH_PAIR=$pr timeout 120 python tools/prof/flash_bench.py >> gpurun_out/${T}_flashbench.log 2>&1; done
  cat gpurun_out/${T}_flashbench.log
fi
