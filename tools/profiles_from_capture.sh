# turn a capture folder (gpurun_out/<tag>, written by tools/capture_r02.sh) into the committed files under profiles/
# usage: bash tools/profiles_from_capture.sh <tag>        (build box: needs cuobjdump / nvdisasm, no GPU)
set -e
cd "$(dirname "$0")/.."
TAG=$1
D=gpurun_out/$TAG
python tools/ncu_raw_summary.py --summary $TAG $D > profiles/r02_ncu_summary.json
cp $D/launches.csv profiles/${TAG}_launches.csv
cp $D/bench_short.json profiles/${TAG}_bench_short.json
for r in c4 mixed full planner planner_acc aroc; do
  [ -f $D/${r}_raw.csv ] && python tools/ncu_raw_summary.py $D/${r}_raw.csv > profiles/${TAG}_${r}_ncu_launches.txt
done
# per-source-region attribution of executed instructions and stall samples
rm -rf /tmp/sass_$TAG && mkdir -p /tmp/sass_$TAG && (cd /tmp/sass_$TAG && cuobjdump -xelf all $OLDPWD/motionplanner_b200/csrc/libmpcheck.so > /dev/null && nvdisasm --print-line-info-inline *.cubin > all.sass 2>/dev/null)
python tools/ncu_regions.py $TAG /tmp/sass_$TAG/all.sass > profiles/${TAG}_k_check_by_region.txt
python tools/sass_extract.py > profiles/r02_sass_extract.txt
ls -la profiles | tail -15
