mkdir -p gpurun_out
T=r02A
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/${T}_bench.json
timeout 420 python tools/prof/drift.py --steps 200 --out gpurun_out/${T}_drift.json > gpurun_out/${T}_drift.log 2>&1; head -8 gpurun_out/${T}_drift.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_step.csv python tools/prof/ncu_step.py > gpurun_out/${T}_ncu_step.log 2>&1
cp gpurun_out/step_ops.json gpurun_out/${T}_step_ops.json
PSTAGE=0 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_step_s0.csv python tools/prof/ncu_step.py > gpurun_out/${T}_ncu_step_s0.log 2>&1
cp gpurun_out/step_ops.json gpurun_out/${T}_step_ops_s0.json
FRIDO_TC_PAIR=2 FRIDO_SK=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair -s 3 -c 1 -o gpurun_out/${T}_pair python tools/prof/conv_bench.py 9 > gpurun_out/${T}_ncu_pair.log 2>&1; tail -2 gpurun_out/${T}_ncu_pair.log
