# Round-end evidence set on one B200: GPU tests, smoke, bench lines, launch list, full ncu captures, sanitizer.
#   bash tools/final_n1.sh TAG
set -x
TAG=${1:-r2z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputest.log 2>&1; tail -3 gpurun_out/${TAG}_gputest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
for wl in config1 config2 config4 config5; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
# launch list of one steady-state frame + full captures
python tools/profile_run.py config3 14 > gpurun_out/${TAG}_config3_times.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s $((25 * 11)) -c 25 --csv --log-file gpurun_out/${TAG}_launches_config3.csv python tools/profile_run.py config3 14 > /dev/null 2>&1
for k in fine_kernel coarse_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 11 -c 1 -f -o gpurun_out/${TAG}_config3_$k python tools/profile_run.py config3 14 > /dev/null 2>&1
done
for wl in config3_nowipe config5; do
  ncu --set full --clock-control none --import-source on -k regex:fine_kernel -s 11 -c 1 -f -o gpurun_out/${TAG}_${wl}_fine_kernel python tools/profile_run.py $wl 14 > gpurun_out/${TAG}_${wl}_times.log 2>&1
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_scene.py > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer scene ok" gpurun_out/${TAG}_sanitizer_$tool.log | head -3
done
python - $TAG <<'PY'
import json, sys
tag = sys.argv[1]
for f in ("bench_n1", "bench_ref", "bench_config1", "bench_config2", "bench_config4", "bench_config5"):
    try:
        d = json.load(open(f"gpurun_out/{tag}_{f}.json"))
        c = d.get("config", {})
        print(f, round(d.get("ms_per_step", 0), 4), round(d.get("value", 0), 1), {k: round(v, 3) for k, v in c.get("stage_ms", {}).items()}, "e2e", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.get("e2e", {}).items() if k != "resident"}, "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
