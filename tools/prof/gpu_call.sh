mkdir -p gpurun_out
T=r02z
for st in 0 1; do
PSTAGE=$st timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s$st.log 2>&1
echo "s$st default: $(grep GRAPH gpurun_out/${T}_perop_s$st.log | cut -c1-60)"
done
FRIDO_SPADE_SPLIT=1 PSTAGE=1 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s1_split.log 2>&1
echo "s1 split: $(grep GRAPH gpurun_out/${T}_perop_s1_split.log | cut -c1-60)"
FRIDO_TC_PAIR=0 PSTAGE=1 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s1_nopair.log 2>&1
echo "s1 plain nopair: $(grep GRAPH gpurun_out/${T}_perop_s1_nopair.log | cut -c1-60)"
timeout -k 5 1200 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
