# development helper: one visit to the GPU box (tests, A/B timings, bench configurations)
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -eq 0 ] || exit 0
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log
for c in c2 c3 c4 c5-16 c5-64 c5-128; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1
  python - $c <<'PY'
import json,sys
c=sys.argv[1]
ls=[x for x in open(f'gpurun_out/bench_{c}.log') if x.startswith('{')]
if not ls: print(c,'NO LINE'); print(open(f'gpurun_out/bench_{c}.log').read()[-1500:]); sys.exit()
d=json.loads(ls[-1]); print(c, d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'e2e16', d.get('e2e_f16_storage',{}).get('value'), 'api', d.get('reference_api',{}).get('value'), d.get('reference_api',{}).get('error'))
PY
done
