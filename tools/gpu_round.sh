# development helper: one visit to the GPU box -- smoke, the GPU tests, the bench records of the round
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 $O/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.log 2>&1; grep "^{" $O/r02_bench_reference.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_20.log 2>&1; grep "^{" $O/r02_bench_20.log | cut -c1-250
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/r02_bench_200.log 2>&1; grep "^{" $O/r02_bench_200.log | cut -c1-250
for c in c3 c4 c5-16 c5-64 c5-128; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_$c.log 2>&1; grep "^{" $O/r02_bench_$c.log | cut -c1-200; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file $O/r02_pk_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; grep "spacing_table\|render_fused_pk" $O/r02_pk_launches.csv | tail -4 | cut -c60-400
timeout 300 python tools/splat_bench.py > $O/r02_splat.txt 2>&1; tail -3 $O/r02_splat.txt
timeout 300 python tools/inpaint_bench.py 1 8 32 148 592 > $O/r02_inpaint_slivers.txt 2>&1; tail -5 $O/r02_inpaint_slivers.txt
INPAINT_BENCH_CASE=ring timeout 300 python tools/inpaint_bench.py 1 8 32 148 592 > $O/r02_inpaint_ring.txt 2>&1; tail -5 $O/r02_inpaint_ring.txt
timeout 300 python tools/sample_bench.py > $O/r02_sample_bench.txt 2>&1; tail -1 $O/r02_sample_bench.txt
timeout 300 python tools/generate_samples.py 512 12 /tmp/gs_host 0 > $O/r02_generate_host.txt 2>&1; tail -1 $O/r02_generate_host.txt | cut -c1-330
timeout 300 python tools/generate_samples.py 512 12 /tmp/gs_dev 128 > $O/r02_generate_device.txt 2>&1; tail -1 $O/r02_generate_device.txt | cut -c1-330
