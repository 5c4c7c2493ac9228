for u in 2 4 8; do for b in 128 256; do
echo -n "unroll $u block $b: "; MPIFLOW_BLEND_UNROLL=$u MPIFLOW_BLEND_BLOCK=$b SAMPLE_BENCH_ONLY=blend python tools/sample_bench.py 384 1280 64 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['blend_us'],1), round(d['blend_GBps']))"
done; done
