set -x
cd $GRAFT_REPO_ROOT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qn_iter_kernel -s 2 -c 1 -o gpurun_out/r02_iter_kernel python tools/ncu_targets.py fused > gpurun_out/ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mapreduce_kernel -s 6 -c 2 -o gpurun_out/r02_stream_trial python tools/ncu_targets.py stream > gpurun_out/ncu_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep
