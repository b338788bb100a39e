timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 80 python tools/op_bench.py --size 3840x2160 --types f32 --ops pow 2>&1 | grep -i "pow"
timeout 80 python tools/op_bench.py --size 3840x2160 --types f32 --ops colorconvert 2>&1 | grep -i "colorconv"
for b in 1 8 4 16; do
  echo "band $b"
  TOUCAN_STACK_BAND=$b TOUCAN_BENCH_CPU=0 TOUCAN_BENCH_E2E_STEPS=1 timeout 120 python bench.py --steps 96 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
done
