mkdir -p gpurun_out
T=r02V
FRIDO_TC_EPI16=2 timeout -k 5 300 python -m pytest tests/test_gpu_tc.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_k.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_k.log
tail -3 gpurun_out/${T}_k.log
if grep -q "rc=0" gpurun_out/${T}_k.log; then
  for e in 0 2; do echo "== EPI16=$e" >> gpurun_out/${T}.log; FRIDO_TC_EPI16=$e LB_SEL=1,3,4,5,7,8,10,11 timeout 150 python tools/prof/lin_bench.py >> gpurun_out/${T}.log 2>&1; done
  for e in 0 2; do echo "== conv EPI16=$e" >> gpurun_out/${T}.log; FRIDO_TC_EPI16=$e FRIDO_SK=0 timeout 100 python tools/prof/conv_bench.py 9 7 >> gpurun_out/${T}.log 2>&1; done
  cat gpurun_out/${T}.log
fi
