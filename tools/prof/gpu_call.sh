mkdir -p gpurun_out
T=r02O
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_pair.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_tc.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tc.log
tail -5 gpurun_out/${T}_tc.log
if grep -q "rc=0" gpurun_out/${T}_tc.log; then
  run() { echo "== $*" >> gpurun_out/${T}.log; env "$@" FRIDO_TC_PAIR=0 FRIDO_SK=0 timeout 120 python tools/prof/conv_bench.py 9 7 8 4 >> gpurun_out/${T}.log 2>&1; }
  run CB_BNS=192 FRIDO_TC_DECOUPLE=0
  run CB_BNS=64,128,192 FRIDO_TC_DECOUPLE=1
  cat gpurun_out/${T}.log
fi
