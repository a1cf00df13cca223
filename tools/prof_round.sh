set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
python tools/power_log.py > gpurun_out/power_log.txt 2>&1
SHAPE=cfg2 python tools/prof_train.py > gpurun_out/p_cfg2.log 2>&1 && SHAPE=cfg2 ncu --metrics $M --clock-control none -s 72 -c 40 --csv --log-file gpurun_out/r02_train_cfg2_kernels.csv python tools/prof_train.py > gpurun_out/n_cfg2.log 2>&1
SHAPE=ml1m python tools/prof_train.py > gpurun_out/p_ml1m.log 2>&1 && SHAPE=ml1m ncu --metrics $M --clock-control none -s 100 -c 80 --csv --log-file gpurun_out/r02_train_ml1m_kernels.csv python tools/prof_train.py > gpurun_out/n_ml1m.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks train > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks train > gpurun_out/ncu_bench.log 2>&1
NROWS=10000000 python tools/prof_topk.py > gpurun_out/p_topk.log 2>&1 && NROWS=10000000 ncu --set full --clock-control none --import-source on -k regex:stream_scores2 -s 3 -c 1 -o gpurun_out/r02_topk_full python tools/prof_topk.py > gpurun_out/n_topk.log 2>&1
tail -3 gpurun_out/*.log gpurun_out/power_log.txt
