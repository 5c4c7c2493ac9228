# development helper: one visit to the GPU box (tests, probes, A/B timings, bench, one ncu capture)
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
timeout 300 python tools/probe_norm.py > gpurun_out/probe_norm.log 2>&1; cat gpurun_out/probe_norm.log
AB_FRAMES=16 AB_STEPS=20 timeout 900 python tools/ab.py "" "MPIFLOW_PK_BAND_MB=0" "MPIFLOW_PK_BAND_MB=6" "MPIFLOW_PK_BAND_MB=11" "MPIFLOW_PK_BAND_MB=32" "MPIFLOW_PK_DBG=0x10000000" "MPIFLOW_PK_DBG=0x10000000 MPIFLOW_PK_BAND_MB=0" "MPIFLOW_PK_SPACING_TABLE=0" > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-700
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fused_pk -s 2 -c 1 -o gpurun_out/r02_pk -f python tools/pk_once.py 16 4 > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
