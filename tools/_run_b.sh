timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 150 python tools/op_bench.py --size 3840x2160 --types f32 --out gpurun_out/op4k_s3.json > gpurun_out/op4k_s3.log 2>&1; grep -i "pow\|colorconv\|rotate" gpurun_out/op4k_s3.log
timeout 200 python tools/op_bench.py --size 7680x4320 --types f16,u8 --out gpurun_out/op8k_s3.json > gpurun_out/op8k_s3.log 2>&1; grep -i "pow\|colorconv\|rotate" gpurun_out/op8k_s3.log
timeout 300 python tools/config_bench.py --out gpurun_out/config_bench_s3.json > gpurun_out/config_bench_s3.log 2>&1; tail -2 gpurun_out/config_bench_s3.log | cut -c1-300
