# scout launch on / off over batch sizes of the C4 workload (bench.py device-resident value); MPC_SCOUTS forces the count
cd $GRAFT_REPO_ROOT
for nq in 4 8 16 32; do for sc in 0 1; do
  MPC_SCOUTS=$sc python bench.py --no-cpu-baseline --steps 200 --lanes 1 --queries $nq > gpurun_out/ts.json 2> gpurun_out/ts.err
  python -c "
import json; d=json.load(open('gpurun_out/ts.json')); r=d['roofline']; print('queries $nq scouts $sc | ms %.4f kernel_ms %.4f' % (d['ms_per_step'], r['kernel_ms']))"
done; done
