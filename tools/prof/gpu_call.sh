mkdir -p gpurun_out
T=r02k
timeout -k 5 240 python -m pytest tests/test_gpu_flash.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_flash.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_flash.log
tail -15 gpurun_out/${T}_flash.log
if grep -q "rc=0" gpurun_out/${T}_flash.log; then
  for f in 0 1; do
    FRIDO_FLASH=$f PSTAGE=0 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s0_f$f.log 2>&1
    echo "flash $f: $(grep GRAPH gpurun_out/${T}_perop_s0_f$f.log | cut -c1-60)"; grep "attn1" gpurun_out/${T}_perop_s0_f$f.log | head -12
  done
  timeout -k 5 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_benched.py -x -q --timeout=200 > gpurun_out/${T}_model.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_model.log
  tail -4 gpurun_out/${T}_model.log
fi
