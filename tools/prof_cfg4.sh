M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
SHAPE=cfg4 python tools/prof_train.py > gpurun_out/p_cfg4.log 2>&1 && SHAPE=cfg4 ncu --metrics $M --clock-control none -s 140 -c 70 --csv --log-file gpurun_out/r02_train_cfg4_kernels.csv python tools/prof_train.py > gpurun_out/n_cfg4.log 2>&1
tail -n 2 gpurun_out/p_cfg4.log gpurun_out/n_cfg4.log
