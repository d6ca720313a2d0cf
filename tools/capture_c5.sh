set -e
cd $GRAFT_REPO_ROOT
python tools/bench_c5.py --iters 3 > gpurun_out/c5_plain.json 2> gpurun_out/c5_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_check --launch-skip 3 --launch-count 1 -o gpurun_out/r01k_c5_k_check python tools/bench_c5.py --iters 3 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out | tail -3
