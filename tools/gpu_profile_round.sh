# The round-2 profile capture (run under gpurun from the repo root; results land in gpurun_out/).
#   1. the default bench line
#   2. ncu launch list of a short bench command (device time per launch; shares, not absolutes)
#   3. ncu --set full (+ the FP64 tensor-core counters) of the hot kernels of that command
#   4. the same for the complex128 engine and for the on-chip run kernel
set -x
R=r02
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench.err || exit 1
CMD="python bench.py --steps 2 --warmup 3 --batch 4096 --sweep-points 0 --e2e-steps 1 --cpu-seconds 1 --energy-steps 2 --converged-networks 0 --c128-batch 0 --e2e-run-networks 0"
EXTRA="sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:"lq_chol|recompose_chol|svd_warp|pair_rdm_dmma" -s 20 -c 4 -o gpurun_out/${R}_full $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
python tools/ncu_summary.py launches gpurun_out/${R}_launches.csv > gpurun_out/${R}_launches.txt
python tools/ncu_summary.py full gpurun_out/${R}_full.ncu-rep > gpurun_out/${R}_ncu_full_summary.txt
python tools/ncu_summary.py edge gpurun_out/${R}_ncu_full_summary.txt 4096 | sed "s#gpurun_out/#profiles/#" > gpurun_out/${R}_ncu_edge_update.json
rm -f gpurun_out/${R}_full.ncu-rep
CMDC="$CMD --dtype c128 --batch 2048"
$CMDC > gpurun_out/plain_c.json 2> gpurun_out/plain_c.err && \
ncu --set full --metrics $EXTRA --clock-control none -k regex:"lq_chol|recompose_chol|svd_kernel" -s 12 -c 3 -o gpurun_out/${R}_full_c128 $CMDC > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
python tools/ncu_summary.py full gpurun_out/${R}_full_c128.ncu-rep > gpurun_out/${R}_ncu_full_summary_c128.txt
rm -f gpurun_out/${R}_full_c128.ncu-rep   # gpurun_out is capped at 64 MiB: only the summaries travel back
python tools/probes/gpu_fused_ncu.py > gpurun_out/plain_f.log 2>&1 && \
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:run_fused -c 1 -o gpurun_out/${R}_full_fused python tools/probes/gpu_fused_ncu.py > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
python tools/ncu_summary.py full gpurun_out/${R}_full_fused.ncu-rep > gpurun_out/${R}_ncu_full_summary_fused.txt
ncu -i gpurun_out/${R}_full_fused.ncu-rep --page source --csv > gpurun_out/${R}_fused_source.csv
rm -f gpurun_out/${R}_full_fused.ncu-rep
ls -la gpurun_out
