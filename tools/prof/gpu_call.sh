mkdir -p gpurun_out
T=r02W
timeout -k 5 1200 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/${T}_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
