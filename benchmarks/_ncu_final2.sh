set -x
timeout 400 ncu --set full --import-source on --clock-control none -k regex:roi_bwd_resize --launch-skip 3 -c 1 -o gpurun_out/roi_bwd_resize_r01c python benchmarks/stages.py --only roi_bwd --modes resize --iters 1 > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:roi_bwd_max --launch-skip 3 -c 1 -o gpurun_out/roi_bwd_max_r01c python benchmarks/stages.py --only roi_bwd --modes max --iters 1 > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:roi_fwd_kernel --launch-skip 3 -c 1 -o gpurun_out/roi_fwd_max_r01c python benchmarks/stages.py --only roi_fwd --modes max --iters 1 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log
