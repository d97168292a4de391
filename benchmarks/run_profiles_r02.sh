#!/bin/bash
# Round-2 evidence pass on one B200 (gpurun --timeout 1800 -- 'bash benchmarks/run_profiles_r02.sh'): launch list of the
# bench step, one `ncu --set full` capture per hot kernel, stage timings.  Outputs land in gpurun_out/; the summaries
# kept under profiles/ are produced from them by benchmarks/ncu_summary.py (no GPU needed).
set -x
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 10 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-seconds 0.05 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 300 $NCU -k regex:roi_fwd_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_roi_fwd_resize_c1 \
    python benchmarks/roi_one.py --dir fwd --mode resize --rois 320 --batch 64 > /dev/null 2>&1
timeout 300 $NCU -k regex:roi_fwd_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_roi_fwd_max_c5x8 \
    python benchmarks/roi_one.py --dir fwd --mode max --rois 2000 --batch 8 > /dev/null 2>&1
timeout 300 $NCU -k regex:roi_bwd --launch-skip 6 -c 3 -o gpurun_out/r02_roi_bwd_resize_c5x8 \
    python benchmarks/roi_one.py --dir bwd --mode resize --rois 2000 --batch 8 > /dev/null 2>&1
timeout 300 $NCU -k regex:roi_bwd --launch-skip 6 -c 3 -o gpurun_out/r02_roi_bwd_max_c5x8 \
    python benchmarks/roi_one.py --dir bwd --mode max --rois 2000 --batch 8 > /dev/null 2>&1
timeout 300 $NCU -k regex:roi_bwd --launch-skip 6 -c 3 -o gpurun_out/r02_roi_bwd_resize_c5x1 \
    python benchmarks/roi_one.py --dir bwd --mode resize --rois 2000 --batch 1 > /dev/null 2>&1
timeout 300 $NCU -k regex:nms_i16 --launch-skip 2 -c 1 -o gpurun_out/r02_nms_train_b1 \
    python benchmarks/prop_one.py 1 12000 2000 > /dev/null 2>&1
timeout 300 $NCU -k regex:proposals_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_proposals_b1 \
    python benchmarks/prop_one.py 1 8000 300 > /dev/null 2>&1
timeout 300 $NCU -k regex:proposals_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_proposals_b64 \
    python benchmarks/prop_one.py 64 8000 300 > /dev/null 2>&1
timeout 600 python benchmarks/stages.py --iters 30 --json gpurun_out/r02_stages.json > gpurun_out/r02_stages.log 2>&1
ls -la gpurun_out/r02_*
