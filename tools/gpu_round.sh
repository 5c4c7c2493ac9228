# development helper: one visit to the GPU box (tests, A/B timings, bench)
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log; grep -n "bench-shape" gpurun_out/pytest.log
export AB_FRAMES=16 AB_STEPS=20
timeout 900 python tools/ab.py "" "" "MPIFLOW_PK_DBG=0x20000000" "MPIFLOW_PK_DBG=0x10000000" "MPIFLOW_PK_SPACING_TABLE=0" > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
AB_FRAMES=1 AB_STEPS=40 timeout 300 python tools/ab.py "" > gpurun_out/ab1.log 2>&1; cat gpurun_out/ab1.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]
d=json.loads(l); print({k:d[k] for k in ('value','e2e','roofline','reference_api','sample_assembly') if k in d})
PY
