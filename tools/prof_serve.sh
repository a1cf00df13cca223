NROWS=50000000 KQ=1000 python tools/k1000.py > gpurun_out/p_k1000.log 2>&1 && NROWS=50000000 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 60 --csv --log-file gpurun_out/r02_k1000_launches.csv python tools/k1000.py > gpurun_out/n_k1000.log 2>&1
cat gpurun_out/p_k1000.log
