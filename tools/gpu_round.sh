# development helper: one visit to the GPU box
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.log 2>&1; grep "^{" gpurun_out/r02_bench_reference.log | cut -c1-700
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_20.log 2>&1; grep "^{" gpurun_out/r02_bench_20.log | cut -c1-250
timeout 900 python bench.py --steps 25 --warmup 5 > gpurun_out/r02_bench_25.log 2>&1; grep "^{" gpurun_out/r02_bench_25.log | cut -c1-250
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/r02_pk_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; grep "spacing_table\|render_fused_pk" gpurun_out/r02_pk_launches.csv | tail -4 | cut -c60-400
