timeout 250 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/op_bench.py --size 7680x4320 --types f16,u8 --ops blur 2>&1 | grep -i "blur"
timeout 100 python tools/op_bench.py --size 7680x4320 --types f16,u8 --ops unsharp 2>&1 | grep -i "unsharp"
timeout 100 python tools/op_bench.py --size 3840x2160 --types f32 --ops blur 2>&1 | grep -i "blur"
