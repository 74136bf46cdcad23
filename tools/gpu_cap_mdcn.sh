mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^mdcn_tcs_kernel$" -s 2 -c 1 -o gpurun_out/prof_mdcn_fwd -f python tools/one_frame.py 1 > gpurun_out/ncu_mdcn_fwd.log 2>&1
ls -la gpurun_out/prof_mdcn_fwd.ncu-rep
