# development helper: one visit to the GPU box
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
for m in 0 1 2; do timeout 120 tests/probes/mbar_litmus $m; done > gpurun_out/mbar_litmus.log 2>&1; cat gpurun_out/mbar_litmus.log
# sanitizers: memcheck on the benchmarked launch shape and the small-shape suites; racecheck on one small render
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_parity_fullsize_gpu.py::test_bench_launch_shape_against_the_unmodified_reference" -m gpu -q -p no:cacheprovider > gpurun_out/memcheck_bench_shape.log 2>&1; echo "memcheck bench-shape rc=$?"; tail -4 gpurun_out/memcheck_bench_shape.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py tests/test_f16_storage_gpu.py tests/test_splat_gpu.py -m gpu -q -p no:cacheprovider -k "not full_size and not properties_at_full and not litmus and not 1280" > gpurun_out/memcheck_small.log 2>&1; echo "memcheck small rc=$?"; tail -4 gpurun_out/memcheck_small.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest "tests/test_render_gpu.py::test_batch_entries_are_independent" -m gpu -q -p no:cacheprovider > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/racecheck.log; grep "RACECHECK SUMMARY" gpurun_out/racecheck.log; grep -A3 "hazard detected" gpurun_out/racecheck.log | grep " at \|hazard" | sort | uniq -c | sort -rn | head -20
