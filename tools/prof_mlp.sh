SHAPE=ml1m python tools/prof_train.py > gpurun_out/p_mlp.log 2>&1 && SHAPE=ml1m ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 60 -c 12 -f -o gpurun_out/r02_mlp_fused python tools/prof_train.py > gpurun_out/n_mlp.log 2>&1
tail -n 2 gpurun_out/n_mlp.log
