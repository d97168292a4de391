set -x
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"nms_i16|topk_kernel" --launch-skip 8 -c 2 -o gpurun_out/prop_c1b64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_p.log 2>&1
tail -n 3 gpurun_out/ncu_p.log
