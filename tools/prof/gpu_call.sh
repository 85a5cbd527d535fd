mkdir -p gpurun_out
T=r02q
timeout -k 5 400 python -m pytest tests/test_gpu_kernels.py -x -q --timeout=120 -k "attn" > gpurun_out/${T}_attn.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_attn.log
tail -3 gpurun_out/${T}_attn.log
for st in 0 1; do
PALL=1 PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s$st.log 2>&1
echo "$(grep GRAPH gpurun_out/${T}_perop_s$st.log | cut -c1-60)"; grep "^attn2.block\|^attn1.fused" gpurun_out/${T}_perop_s$st.log
done
grep "attn2.block  \|attn1.fused  " gpurun_out/${T}_perop_s0.log | sort | uniq -c | sort -rn | head -12
