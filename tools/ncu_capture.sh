# ncu captures of round 2 (run on a GPU box from the repo root: bash tools/ncu_capture.sh; outputs in gpurun_out/).
# Every profiled command first runs once WITHOUT ncu; numbers printed under ncu are never bench values.
set -x
cd ${GRAFT_REPO_ROOT:-.}
for t in sym fused stream syrk batched; do timeout 200 python tools/ncu_targets.py $t > gpurun_out/ncu_plain_$t.log 2>&1 || exit 1; done
# launch list of the default bench command (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench.log 2>&1
# full captures of the dominant kernels
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qn_lazy_sym_kernel -s 6 -c 2 -o gpurun_out/r02_sym_pass python tools/ncu_targets.py sym > gpurun_out/ncu_sym.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qn_iter_kernel -s 2 -c 1 -o gpurun_out/r02_iter_kernel python tools/ncu_targets.py fused > gpurun_out/ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mapreduce_kernel -s 6 -c 2 -o gpurun_out/r02_stream_trial python tools/ncu_targets.py stream > gpurun_out/ncu_stream.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:syrk_dmma -s 1 -c 1 -o gpurun_out/r02_syrk_48x2 python tools/ncu_targets.py syrk > gpurun_out/ncu_syrk.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:batched_bfgs -c 1 -o gpurun_out/r02_batched python tools/ncu_targets.py batched > gpurun_out/ncu_batched.log 2>&1
ls -la gpurun_out/*.ncu-rep
