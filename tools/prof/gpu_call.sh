mkdir -p gpurun_out
T=r02o
for pdl in 0 1; do
  for c in l2i_coco sg2i_vg; do
    FRIDO_PDL=$pdl PCONFIG=$c PSTAGE=1 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_${c}_pdl$pdl.log 2>&1
    echo "$c pdl=$pdl: $(grep GRAPH gpurun_out/${T}_perop_${c}_pdl$pdl.log | cut -c1-70)"
  done
done
