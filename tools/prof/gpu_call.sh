mkdir -p gpurun_out
T=r02S
FRIDO_SPADE_SPLIT=1 PSTAGE=1 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s1_split.log 2>&1
echo "s1 split: $(grep GRAPH gpurun_out/${T}_perop_s1_split.log | cut -c1-60)"
FRIDO_TC_PAIR=1 PSTAGE=1 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s1_pair.log 2>&1
echo "s1 pair=1: $(grep GRAPH gpurun_out/${T}_perop_s1_pair.log | cut -c1-60)"
PSTAGE=1 timeout 200 python tools/prof/perop.py > gpurun_out/${T}_perop_s1.log 2>&1
echo "s1 default: $(grep GRAPH gpurun_out/${T}_perop_s1.log | cut -c1-60)"
timeout -k 5 1200 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
head -c 700 gpurun_out/${T}_bench.json
