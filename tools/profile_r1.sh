# Round-1 profile set: launch list of one steady-state frame + full captures of the heaviest kernels (config3).
set -x
mkdir -p gpurun_out
TAG=${1:-r1b}
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 31 --csv --log-file gpurun_out/${TAG}_launches_config3.csv python tools/profile_run.py config3 14 > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
for k in coarse_kernel fine_kernel flatten_subdivide_kernel flatten_eseg_emit_kernel path_count_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 11 -c 1 -f -o gpurun_out/${TAG}_$k python tools/profile_run.py config3 14 > /dev/null 2>&1
done
ls -la gpurun_out | tail -8
