timeout 300 python -m pytest tests -m gpu -q -x -k "narrow or batch_matches or streaming" 2>&1 | tail -5
cat gpurun_out/test_notes.txt | tail -3
Q="--steps 4 --warmup 3 --sweep-points 256 --converged-networks 0 --c128-batch 0 --e2e-run-networks 0 --e2e-steps 1 --cpu-seconds 1 --energy-steps 1 --sweep-cpu-points 0"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"]), d["sweep"]["value"], d["sweep_large"]["value"], d["sweep_large"]["seconds"])'
timeout 300 python bench.py $Q 2>gpurun_out/err.log | python -c "$pick" || tail -5 gpurun_out/err.log
