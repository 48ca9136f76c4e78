#!/bin/bash
# Round-2 session m (not a test): compute-sanitizer memcheck / racecheck / synccheck over every kernel of the path.
mkdir -p gpurun_out
timeout 60 python tests/sanitize_run.py > gpurun_out/r2m_plain.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_plain.log
for tool in memcheck racecheck synccheck; do
  ( time timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tests/sanitize_run.py ) > gpurun_out/r2m_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2m_$tool.log
done
for f in gpurun_out/r2m_*.log; do echo "== $f"; grep -c "^ok" $f; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|real" $f | tail -4; done
