set -x
# Round profile: launch list of the bench command + one `ncu --set full` capture of every dominant kernel (each only
# after the same command has exited 0 without ncu).  Outputs in gpurun_out/r1e_*; summaries: tools/ncu_summary.py.
timeout 300 python bench.py --steps 2 --warmup 1 --no-batch4 > gpurun_out/r1e_bench_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1e_launches.csv python bench.py --steps 2 --warmup 1 --no-batch4 > gpurun_out/r1e_ncu0.log 2>&1
timeout 100 python tools/one_step.py 2 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_fwd_chain --launch-skip 1 --launch-count 1 -o gpurun_out/r1e_chain python tools/one_step.py 2 > gpurun_out/r1e_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:block_(bwd_pre|bwd_dx)_umma' --launch-skip 100 --launch-count 4 -o gpurun_out/r1e_blocks python tools/one_step.py 2 > gpurun_out/r1e_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_wgrad_all --launch-skip 1 --launch-count 1 -o gpurun_out/r1e_wgrad python tools/one_step.py 2 > gpurun_out/r1e_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_kernel --launch-skip 9 --launch-count 9 -o gpurun_out/r1e_gemm python tools/one_step.py 2 > gpurun_out/r1e_ncu3.log 2>&1
for f in gpurun_out/r1e_ncu1.log gpurun_out/r1e_ncu2.log gpurun_out/r1e_ncu3.log; do tail -n 2 $f; done
