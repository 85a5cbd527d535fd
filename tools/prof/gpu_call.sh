mkdir -p gpurun_out
T=r02i
timeout -k 5 150 python -m pytest tests/test_gpu_nf.py -x -q --timeout=60 -p no:cacheprovider > gpurun_out/${T}_nf.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_nf.log
tail -3 gpurun_out/${T}_nf.log
if grep -q "rc=0" gpurun_out/${T}_nf.log; then
  timeout -k 5 300 python -m pytest tests/test_gpu_model.py -x -q --timeout=150 -k "norm_modes or fusion_switches" > gpurun_out/${T}_model.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_model.log
  tail -4 gpurun_out/${T}_model.log
  for st in 0 1; do for m in 0 1 2 auto; do
    FRIDO_FUSE_NORM=$m PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s${st}_m${m}.log 2>&1
    echo "stage $st mode $m: $(grep GRAPH gpurun_out/${T}_perop_s${st}_m${m}.log | cut -c1-60)"
  done; done
fi
