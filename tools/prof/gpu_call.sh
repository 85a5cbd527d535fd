mkdir -p gpurun_out
T=r02C
for cfg in "1 0" "0 2" "1 1"; do set -- $cfg; echo "SK=$1 PAIR=$2" >> gpurun_out/${T}_convbench.log
FRIDO_SK=$1 FRIDO_TC_PAIR=$2 timeout 150 python tools/prof/conv_bench.py 7 8 9 4 5 0 1 >> gpurun_out/${T}_convbench.log 2>&1; done
cat gpurun_out/${T}_convbench.log
for cfg in "1 0" "0 2" "1 1"; do set -- $cfg; echo "SK=$1 PAIR=$2" >> gpurun_out/${T}_linbench.log
FRIDO_SK=$1 FRIDO_TC_PAIR=$2 LB_SEL=1,3,4,5,7,8,10 timeout 150 python tools/prof/lin_bench.py >> gpurun_out/${T}_linbench.log 2>&1; done
cat gpurun_out/${T}_linbench.log
