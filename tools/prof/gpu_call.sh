mkdir -p gpurun_out
T=r02l
timeout -k 5 240 python -m pytest tests/test_gpu_flash.py tests/test_gpu_pack.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_flash.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_flash.log
tail -8 gpurun_out/${T}_flash.log
if grep -q "rc=0" gpurun_out/${T}_flash.log; then
  timeout 120 python tools/prof/flash_bench.py > gpurun_out/${T}_flashbench.log 2>&1; cat gpurun_out/${T}_flashbench.log
  PSTAGE=0 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s0.log 2>&1
  echo "$(grep GRAPH gpurun_out/${T}_perop_s0.log | cut -c1-60)"; grep "attn1" gpurun_out/${T}_perop_s0.log | head -6
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 3 -c 1 -o gpurun_out/${T}_flash python tools/prof/flash_bench.py 16x1024x384 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
  timeout -k 5 900 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
  tail -4 gpurun_out/${T}_pytest.log
fi
