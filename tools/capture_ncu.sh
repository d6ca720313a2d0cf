# launch list of the bench command + one ncu --set full capture of the two k_check launches (scout, main) of one step
set -e
cd $GRAFT_REPO_ROOT
TAG=${1:-r01n}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/cap_plain.json 2>gpurun_out/cap_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/ncu1.log 2>&1
if [ "$2" = "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_check --launch-skip 4 --launch-count 2 -o gpurun_out/${TAG}_k_check python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/ncu2.log 2>&1
fi
ls -la gpurun_out
