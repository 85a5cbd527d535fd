mkdir -p gpurun_out
T=r02Y
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py -x -q --timeout=60 -p no:cacheprovider -k "variants" > gpurun_out/${T}_k.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_k.log
tail -3 gpurun_out/${T}_k.log
