# groups per CTA of the main launch x groups per CTA of the scout launch on the C4 batch (bench.py device-resident value)
cd $GRAFT_REPO_ROOT
for cfg in "64 16" "128 16" "64 8" "64 32" "128 8" "32 16"; do
  set -- $cfg
  MPC_GPC=$1 MPC_SCOUT_GPC=$2 python bench.py --no-cpu-baseline --steps 300 --lanes 1 > gpurun_out/ts.json 2> gpurun_out/ts.err || tail -3 gpurun_out/ts.err
  python -c "
import json; d=json.load(open('gpurun_out/ts.json')); r=d['roofline']; print('gpc $1 scout gpc $2 | value %.3e ms %.4f kernel_ms %.4f | looked at %.4f | warm %.4f | one query %.4f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['executed'].get('pairs_looked_at_frac', -1), d['warm_l2']['ms_per_step'], d['latency']['device_ms_single_query']))"
done
python tools/bench_c5.py --lanes 1 --check 2>/dev/null | cut -c1-420
MPC_SLOT_ORDER=2 python tools/bench_c5.py --lanes 1 2>/dev/null | cut -c1-420
