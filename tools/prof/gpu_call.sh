mkdir -p gpurun_out
T=r02E
timeout -k 5 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nf.py -x -q --timeout=120 -k "norm or gn or nf or finalize" > gpurun_out/${T}_k.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_k.log; tail -3 gpurun_out/${T}_k.log
for st in 0 1; do
PALL=1 PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s$st.log 2>&1
echo "s$st: $(grep GRAPH gpurun_out/${T}_perop_s$st.log | cut -c1-60)"; grep "^res.norm\|^st.norm\|^gn_finalize" gpurun_out/${T}_perop_s$st.log
done
timeout -k 5 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_benched.py -x -q --timeout=300 > gpurun_out/${T}_model.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_model.log; tail -3 gpurun_out/${T}_model.log
