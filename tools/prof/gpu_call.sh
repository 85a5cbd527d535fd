mkdir -p gpurun_out
T=r02AA
timeout 420 python tools/prof/drift.py --steps 200 --out gpurun_out/${T}_drift.json > gpurun_out/${T}_drift.log 2>&1; head -12 gpurun_out/${T}_drift.json
