set -x
# Round profile (R = round tag, default r2): launch list of the bench command + one `ncu --set full` capture of every
# dominant kernel (each only after the same command has exited 0 without ncu).  Outputs in gpurun_out/${R}_*; summaries
# with tools/ncu_summary.py (launches / full / traffic) go to profiles/.
R=${R:-r2}
timeout 300 python bench.py --steps 2 --warmup 3 --no-batch4 --no-fastgen --no-cpu-baseline > gpurun_out/${R}_bench_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-batch4 --no-fastgen --no-cpu-baseline > gpurun_out/${R}_ncu0.log 2>&1
timeout 100 python tools/one_step.py 2 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_fwd_chain --launch-skip 1 --launch-count 1 -o gpurun_out/${R}_fwd_chain python tools/one_step.py 2 > gpurun_out/${R}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_bwd_chain_f --launch-skip 1 --launch-count 1 -o gpurun_out/${R}_bwd_chain python tools/one_step.py 2 > gpurun_out/${R}_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block_bwd_fused_reduce|post2_xent" --launch-skip 2 --launch-count 2 -o gpurun_out/${R}_wgrad python tools/one_step.py 2 > gpurun_out/${R}_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_kernel --launch-skip 8 --launch-count 8 -o gpurun_out/${R}_gemm python tools/one_step.py 2 > gpurun_out/${R}_ncu3.log 2>&1
timeout 100 python tools/one_step_cfg5.py 2 > /dev/null 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_cfg5_launches.csv python tools/one_step_cfg5.py 2 > gpurun_out/${R}_ncu_cfg5.log 2>&1
for f in gpurun_out/${R}_ncu1.log gpurun_out/${R}_ncu2.log gpurun_out/${R}_ncu3.log gpurun_out/${R}_ncu4.log; do tail -n 2 $f; done
