mkdir -p gpurun_out
T=r02G
for dbg in 0 1; do
  if [ $dbg = 1 ]; then export FRIDO_TC_DBG_BULKW=1; fi
  echo "bulkW=$dbg" >> gpurun_out/${T}_convbench.log
  FRIDO_TC_PAIR=2 FRIDO_SK=0 timeout 120 python tools/prof/conv_bench.py 7 8 9 >> gpurun_out/${T}_convbench.log 2>&1
done
cat gpurun_out/${T}_convbench.log
