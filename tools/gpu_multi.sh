# development helper: multi-GPU bench lines (N = $1), incl. the output-gather leg over the transports given after N
set -x
cd $GRAFT_REPO_ROOT
N=$1; shift
MODES=${@:-p2p push nccl}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@"; }
for g in $MODES; do
  run --gather $g > gpurun_out/bench_n${N}_$g.log 2>&1
  python - $N $g <<'PY'
import json,sys
n,r=sys.argv[1:3]
ls=[x for x in open(f'gpurun_out/bench_n{n}_{r}.log') if x.startswith('{')]
if not ls: print(n,r,'NO LINE'); print(open(f'gpurun_out/bench_n{n}_{r}.log').read()[-3000:]); sys.exit()
d=json.loads(ls[-1]); g=d.get('with_output_gather') or {}
print('N',n,'gather',r,'value',d['value'],'with gather',g.get('value'),g.get('transport'),g.get('verified'),'e2e',d['e2e']['value'],'e2e16',d['e2e_f16_storage']['value'])
PY
done
