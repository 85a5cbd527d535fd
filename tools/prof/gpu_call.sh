mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 420 python tools/prof/drift.py --steps 200 --out gpurun_out/r02a_drift.json > gpurun_out/r02a_drift.log 2>&1
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 2 -o gpurun_out/r02a_conv_long python tools/prof/conv_bench.py 9 7 > gpurun_out/r02a_ncu_long.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02a_step.csv python tools/prof/ncu_step.py > gpurun_out/r02a_ncu_step.log 2>&1
cp gpurun_out/step_ops.json gpurun_out/r02a_step_ops.json
tail -3 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_drift.json; cat gpurun_out/r02a_bench.json | head -c 3000
