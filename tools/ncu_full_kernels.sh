mkdir -p gpurun_out/r2n
CASES=fc1,proj,qkv,fc2dg,attn,ln,wgrad,blend,optim,loss timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gemm|attn16|layernorm|clip_adamw|sqnorm|distill' -o /tmp/kern_a python tests/bringup/kern_cases.py > gpurun_out/r2n/ncu_a.log 2>&1; echo "ncu a rc=$?"; grep -c Profiling gpurun_out/r2n/ncu_a.log
ncu -i /tmp/kern_a.ncu-rep --page raw --csv > gpurun_out/r2n/kern_a_raw.csv 2>/dev/null; ls -la /tmp/kern_a.ncu-rep gpurun_out/r2n/kern_a_raw.csv
CASES=t2t timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'performer|unfold' -o /tmp/kern_b python tests/bringup/kern_cases.py > gpurun_out/r2n/ncu_b.log 2>&1; echo "ncu b rc=$?"; grep -c Profiling gpurun_out/r2n/ncu_b.log
ncu -i /tmp/kern_b.ncu-rep --page raw --csv > gpurun_out/r2n/kern_b_raw.csv 2>/dev/null; ls -la gpurun_out/r2n/
