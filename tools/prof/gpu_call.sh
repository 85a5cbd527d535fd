mkdir -p gpurun_out
T=r02R
timeout -k 5 400 python -m pytest tests/test_gpu_nf.py tests/test_gpu_flash.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_k.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_k.log
tail -3 gpurun_out/${T}_k.log
if grep -q "rc=0" gpurun_out/${T}_k.log; then
  echo "== nf" >> gpurun_out/${T}.log; FRIDO_SK=0 NB_ONLY=nf timeout 120 python tools/prof/nf_bench.py 0 1 2 3 >> gpurun_out/${T}.log 2>&1
  echo "== flash" >> gpurun_out/${T}.log; timeout 120 python tools/prof/flash_bench.py >> gpurun_out/${T}.log 2>&1
  cat gpurun_out/${T}.log
  for st in 0 1; do
    PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s$st.log 2>&1
    echo "s$st: $(grep GRAPH gpurun_out/${T}_perop_s$st.log | cut -c1-60)"
  done
fi
