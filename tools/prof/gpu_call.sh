mkdir -p gpurun_out
T=r02Q
timeout -k 5 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_pair.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_k.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_k.log
tail -3 gpurun_out/${T}_k.log
if grep -q "rc=0" gpurun_out/${T}_k.log; then
  echo "== single decoupled" >> gpurun_out/${T}.log; CB_BNS=64,192 FRIDO_TC_PAIR=0 FRIDO_SK=0 timeout 120 python tools/prof/conv_bench.py 9 7 8 4 >> gpurun_out/${T}.log 2>&1
  echo "== lin" >> gpurun_out/${T}.log; LB_SEL=1,3,4,5,7,8,10 timeout 150 python tools/prof/lin_bench.py >> gpurun_out/${T}.log 2>&1
  cat gpurun_out/${T}.log
  for st in 0 1; do
    PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s$st.log 2>&1
    echo "s$st: $(grep GRAPH gpurun_out/${T}_perop_s$st.log | cut -c1-60)"
  done
fi
