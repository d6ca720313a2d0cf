set -e
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/capl_plain.json 2>gpurun_out/capl_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_check --launch-skip 4 --launch-count 1 -o gpurun_out/r01l_k_check python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
