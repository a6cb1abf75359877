# Round-2 profile set: launch list of one steady-state frame (config3) + full captures of the heaviest kernels.
#   bash tools/profile_r2.sh TAG "kernel list" [workload] [launches per frame]
set -x
mkdir -p gpurun_out
TAG=${1:-r2a}
KERNELS=${2:-"fine_kernel coarse_kernel"}
WL=${3:-config3}
NL=${4:-31}
# the first render takes ~6 growth passes; skip well past them, then capture exactly one frame
python tools/profile_run.py $WL 14 > gpurun_out/${TAG}_${WL}_times.log 2>&1
tail -3 gpurun_out/${TAG}_${WL}_times.log
ncu --metrics gpu__time_duration.sum --clock-control none -s $((NL * 11)) -c $NL --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv python tools/profile_run.py $WL 14 > /dev/null 2>&1
for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 11 -c 1 -f -o gpurun_out/${TAG}_${WL}_$k python tools/profile_run.py $WL 14 > /dev/null 2>&1
done
ls -la gpurun_out | tail -6
