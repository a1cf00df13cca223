python tools/prof_inbatch.py > gpurun_out/p_ig.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:inbatch_grad -s 2 -c 1 -o gpurun_out/r02_inbatch_grad python tools/prof_inbatch.py > gpurun_out/n_ig.log 2>&1
python tools/time_inbatch.py > gpurun_out/time_inbatch.log 2>&1
cat gpurun_out/p_ig.log gpurun_out/time_inbatch.log; tail -n 3 gpurun_out/n_ig.log
