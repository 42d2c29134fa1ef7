#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py: memcheck (full), synccheck and racecheck (quick shapes)
mkdir -p gpurun_out
python tools/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?"; tail -2 gpurun_out/sanitize_plain.log
for tool in memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|SANITIZE RUN OK" gpurun_out/sanitize_$tool.log | tail -2
done
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 40 python tools/sanitize_run.py quick > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|SANITIZE RUN OK" gpurun_out/sanitize_racecheck.log | tail -2
