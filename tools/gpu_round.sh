# development helper: one visit to the GPU box (tests, A/B timings of library variants, bench)
set -x
cd $GRAFT_REPO_ROOT
P=$GRAFT_REPO_ROOT/mpi-flow_b200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log; grep -n "bench-shape" gpurun_out/pytest.log
export AB_FRAMES=16 AB_STEPS=20
timeout 900 python tools/ab.py "" "MPIFLOW_PK_DBG=0x20000000" "MPIFLOW_LIB=$P/libmpiflow_sm100_scalar.so" "MPIFLOW_LIB=$P/libmpiflow_sm100_scalar.so MPIFLOW_PK_DBG=0x20000000" "MPIFLOW_LIB=$P/libmpiflow_sm100_ab.so MPIFLOW_PK_CFG=1" "MPIFLOW_LIB=$P/libmpiflow_sm100_ab.so MPIFLOW_PK_CFG=2" > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
MPIFLOW_LIB=$P/libmpiflow_sm100_scalar.so timeout 1200 python -m pytest tests/test_parity_fullsize_gpu.py tests/test_render_gpu.py -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_scalar.log 2>&1; echo "pytest scalar rc=$?"; tail -5 gpurun_out/pytest_scalar.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
