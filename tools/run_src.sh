mkdir -p gpurun_out/r2src
CASES=fc1 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'gemm2' -o /tmp/kern_s python tests/bringup/kern_cases.py > gpurun_out/r2src/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/kern_s.ncu-rep --page source --csv > gpurun_out/r2src/src_fc1.csv 2>/dev/null; ls -la gpurun_out/r2src
