mkdir -p gpurun_out/r2san
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r2san/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|ok|Invalid|Error" gpurun_out/r2san/memcheck.log | head -12
CASES=t2t,compact timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r2san/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ok|hazard|Error" gpurun_out/r2san/racecheck.log | head -12
