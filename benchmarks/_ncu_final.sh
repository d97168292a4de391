set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"roi_(fwd|bwd)" --launch-skip 12 -c 4 -o gpurun_out/roi_c1_r01c python benchmarks/stages.py --only roi --iters 1 > gpurun_out/ncu_roi_c.log 2>&1
tail -3 gpurun_out/ncu_roi_c.log
