#!/bin/bash
# run the bench with every library under build_variants/ swapped in (GPU box scratch copy only)
cd $GRAFT_REPO_ROOT
cp motionplanner_b200/csrc/libmpcheck.so /tmp/base.so
for v in /tmp/base.so build_variants/*.so; do
  cp $v motionplanner_b200/csrc/libmpcheck.so
  python bench.py --steps 300 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v', 'value %.3e e2e %.3e ms %.4f kernel_ms %.4f fulleval %.3e lat %.1f us' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['full_evaluation']['checks_per_s'], 1e3*d['latency']['planning_step_p50_ms']))"
done
cp /tmp/base.so motionplanner_b200/csrc/libmpcheck.so
