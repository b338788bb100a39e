timeout 280 ncu --set full --clock-control none --import-source on -k regex:k_rotate_tile -s 3 -c 1 -o gpurun_out/s3_rotate -f python tools/op_bench.py --size 3840x2160 --types f32 --ops rotate > gpurun_out/s3_rotate_ncu.log 2>&1
tail -3 gpurun_out/s3_rotate_ncu.log
