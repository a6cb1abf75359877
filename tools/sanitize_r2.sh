# compute-sanitizer over the small scenes of tools/sanitizer_scene.py; logs under gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_scene.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer scene ok|hazard" gpurun_out/r2_sanitizer_$tool.log | head -5
done
