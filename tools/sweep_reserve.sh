for r in 16 32 48; do
  IMDBN_CD_SMALL=1 timeout 200 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --pipeline-reserve $r > gpurun_out/sw_cds_$r.json 2> gpurun_out/sw_cds_$r.err
done
for r in 24 32 48; do
  timeout 200 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --pipeline-reserve $r > gpurun_out/sw_tc_$r.json 2> gpurun_out/sw_tc_$r.err
done
timeout 200 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --pipeline-reserve -1 > gpurun_out/sw_nopipe.json 2> gpurun_out/sw_nopipe.err
IMDBN_CD_SMALL=1 timeout 200 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --pipeline-reserve -1 > gpurun_out/sw_nopipe_cds.json 2> gpurun_out/sw_nopipe_cds.err
