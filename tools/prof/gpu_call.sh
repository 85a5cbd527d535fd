mkdir -p gpurun_out
T=r02n
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${T}_bench.json
timeout 420 python tools/prof/drift.py --steps 200 --out gpurun_out/${T}_drift.json > gpurun_out/${T}_drift.log 2>&1; cat gpurun_out/${T}_drift.json | head -30
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_step.csv python tools/prof/ncu_step.py > gpurun_out/${T}_ncu_step.log 2>&1
cp gpurun_out/step_ops.json gpurun_out/${T}_step_ops.json
timeout 500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"; tail -c 800 gpurun_out/${T}_ref.json
