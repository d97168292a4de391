#!/bin/bash
# One GPU-box pass over everything that is measured: GPU tests, per-stage timings, the bench line and the ncu evidence.
#   gpurun --timeout 1800 -- 'bash benchmarks/run_all.sh'      (outputs land in gpurun_out/, copy what is kept to profiles/)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python benchmarks/stages.py --iters 30 --json gpurun_out/stages.json > gpurun_out/stages.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
if [ "$1" = "--ncu" ]; then
  # launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  # one full capture per RoI kernel
  for spec in "roi_fwd_kernel roi_fwd resize" "roi_fwd_kernel roi_fwd max" "roi_bwd_resize roi_bwd resize" "roi_bwd_max roi_bwd max"; do
    set -- $spec
    timeout 400 ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip 3 -c 1 \
        -o gpurun_out/$2_$3 python benchmarks/stages.py --only $2 --modes $3 --iters 1 > gpurun_out/ncu_$2_$3.log 2>&1
  done
fi
