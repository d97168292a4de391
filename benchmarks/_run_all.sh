set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python benchmarks/stages.py --iters 30 --json gpurun_out/stages_r01c.json > gpurun_out/stages_r01c.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
cat gpurun_out/bench_r01c.json
