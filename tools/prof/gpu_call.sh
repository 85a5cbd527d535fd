mkdir -p gpurun_out
T=r02T
FRIDO_TC_PAIR=0 FRIDO_SK=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_bf -s 3 -c 1 -o gpurun_out/${T}_conv python tools/prof/conv_bench.py 9 > gpurun_out/${T}_ncu_conv.log 2>&1; tail -2 gpurun_out/${T}_ncu_conv.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_nf -s 4 -c 1 -o gpurun_out/${T}_nf env FRIDO_SK=0 NB_ONLY=nf python tools/prof/nf_bench.py 0 > gpurun_out/${T}_ncu_nf.log 2>&1; tail -2 gpurun_out/${T}_ncu_nf.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 3 -c 1 -o gpurun_out/${T}_flash python tools/prof/flash_bench.py 16x1024x384 > gpurun_out/${T}_ncu_flash.log 2>&1; tail -2 gpurun_out/${T}_ncu_flash.log
for st in 1 0; do
PSTAGE=$st timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_step_s$st.csv python tools/prof/ncu_step.py > gpurun_out/${T}_ncu_step_s$st.log 2>&1
cp gpurun_out/step_ops.json gpurun_out/${T}_step_ops_s$st.json
done
