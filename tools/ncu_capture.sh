set -x
cd $GRAFT_REPO_ROOT
for t in sym fused stream syrk; do timeout 200 python tools/ncu_targets.py $t > gpurun_out/ncu_plain_$t.log 2>&1 || exit 1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qn_lazy_sym_kernel -s 6 -c 2 -o gpurun_out/r02_sym_pass python tools/ncu_targets.py sym > gpurun_out/ncu_sym.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qn_iter_kernel -s 1 -c 1 -o gpurun_out/r02_iter_kernel python tools/ncu_targets.py fused > gpurun_out/ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stream_trial -s 4 -c 2 -o gpurun_out/r02_stream_trial python tools/ncu_targets.py stream > gpurun_out/ncu_stream.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:syrk_dmma -s 1 -c 1 -o gpurun_out/r02_syrk python tools/ncu_targets.py syrk > gpurun_out/ncu_syrk.log 2>&1
ls -la gpurun_out/*.ncu-rep
