mkdir -p gpurun_out
T=r02I
run() { echo "== $*" >> gpurun_out/${T}.log; env "$@" FRIDO_TC_PAIR=0 FRIDO_SK=0 timeout 120 python tools/prof/conv_bench.py 9 >> gpurun_out/${T}.log 2>&1; }
run CB_BNS=64,192
run CB_BNS=64,192 FRIDO_TC_STAGES=2
run CB_BNS=64,192 FRIDO_TC_STAGES=3
run CB_BNS=64,192 FRIDO_TC_DBG=1
run CB_BNS=64,192 FRIDO_TC_DBG=2
run CB_BNS=64,192 FRIDO_TC_DBG=4
run CB_BNS=64,192 FRIDO_TC_DBG=5
cat gpurun_out/${T}.log
