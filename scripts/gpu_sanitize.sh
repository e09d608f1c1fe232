# NOTE: compute-sanitizer is closed on the gpurun pool of this round (the call is refused); the script is kept for boxes where it is allowed.
mkdir -p gpurun_out
timeout 600 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; tail -2 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all paths ran|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
timeout 900 python -m pytest tests -m gpu -x -q -W ignore -k "distribution_level" 2>&1 | tail -5
