mkdir -p gpurun_out
T=r02AD
timeout -k 5 500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nf.py tests/test_gpu_benched.py -x -q --timeout=300 > gpurun_out/${T}_k.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_k.log
tail -3 gpurun_out/${T}_k.log
timeout -k 5 300 python -m pytest tests/test_gpu_model.py -x -q --timeout=300 -k "golden or full_size or fusion" > gpurun_out/${T}_m.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_m.log
tail -3 gpurun_out/${T}_m.log
