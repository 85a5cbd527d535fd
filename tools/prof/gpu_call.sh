mkdir -p gpurun_out
T=r02X
for c in t2i_clip sg2i_vg l2i_512; do
  FRIDO_BENCH_GPU_EAGER=0 timeout -k 5 700 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
  echo "$c rc=$?"; head -c 300 gpurun_out/${T}_bench_$c.json | tail -c 200; echo
done
