# ncu evidence of one build (run under gpurun): launch list of a short bench command + `--set full` captures of the
# k_check launches of every regime.  usage: bash tools/capture_r02.sh <tag> [regimes...]
set -e
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
shift || true
REG=${@:-c4 mixed full planner}
mkdir -p gpurun_out/$TAG
B="python bench.py --steps 1 --warmup 3 --batches-per-step 8 --no-extras --no-cpu-baseline"
$B > gpurun_out/$TAG/bench_short.json 2> gpurun_out/$TAG/bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/$TAG/launches.csv $B > gpurun_out/$TAG/ncu_launches.log 2>&1
for r in $REG; do
  python tools/ncu_target.py $r 2 > gpurun_out/$TAG/target_$r.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_check|k_refine' -c 6 -o gpurun_out/$TAG/$r python tools/ncu_target.py $r 2 > gpurun_out/$TAG/ncu_$r.log 2>&1
  # only the text pages travel back (gpurun merges at most 64 MiB): raw metrics of all captured launches, source page gzipped
  ncu -i gpurun_out/$TAG/$r.ncu-rep --page raw --csv > gpurun_out/$TAG/${r}_raw.csv
  ncu -i gpurun_out/$TAG/$r.ncu-rep --page source --csv | gzip > gpurun_out/$TAG/${r}_source.csv.gz
  rm -f gpurun_out/$TAG/$r.ncu-rep
done
du -sh gpurun_out/$TAG
ls -la gpurun_out/$TAG
