set -x
python bench.py > gpurun_out/r01j_bench_1gpu.json 2> gpurun_out/r01j_bench.err || exit 1


python bench.py --steps 2 --warmup 3 --batch 4096 --sweep-points 0 --e2e-steps 1 --cpu-seconds 1 --energy-steps 2 > gpurun_out/plain.json 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches.csv python bench.py --steps 2 --warmup 3 --batch 4096 --sweep-points 0 --e2e-steps 1 --cpu-seconds 1 --energy-steps 2 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lq_chol|recompose_chol|svd_warp|pair_rdm_dmma" -s 20 -c 4 -o gpurun_out/r01j_full python bench.py --steps 2 --warmup 3 --batch 4096 --sweep-points 0 --e2e-steps 1 --cpu-seconds 1 --energy-steps 2 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
