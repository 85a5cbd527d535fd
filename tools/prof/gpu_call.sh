mkdir -p gpurun_out
T=r02AE
for v in 0 1; do for st in 0 1; do
FRIDO_FUSE_NORM_1X1=$v PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s${st}_1x1_$v.log 2>&1
echo "1x1=$v s$st: $(grep GRAPH gpurun_out/${T}_perop_s${st}_1x1_$v.log | cut -c1-60)"
done; done
