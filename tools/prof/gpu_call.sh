mkdir -p gpurun_out
T=r02p
run() { # name, env...
  name=$1; shift
  env "$@" PSTAGE=0 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s0_$name.log 2>&1
  echo "$name: $(grep GRAPH gpurun_out/${T}_perop_s0_$name.log | cut -c1-60)"
}
run narrow FRIDO_FUSE_NORM_WIDE=0
run wide FRIDO_FUSE_NORM_WIDE=1
run wide1x1 FRIDO_FUSE_NORM_WIDE=1 FRIDO_FUSE_NORM_1X1=1
run plain FRIDO_FUSE_NORM=0
timeout -k 5 500 python -m pytest tests/test_gpu_model.py -x -q --timeout=300 -k "norm_modes or fusion_switches" > gpurun_out/${T}_model.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_model.log
tail -3 gpurun_out/${T}_model.log
