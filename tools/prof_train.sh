# Per-kernel ncu table of ~one training step: `SHAPE=cfg2 START=76 COUNT=46 OUT=r02_train_cfg2_kernels_v3 bash tools/prof_train.sh`
# (cfg2: START 70-76 COUNT 40-46; ml1m: START 100 COUNT 60-80; cfg4: START 140 COUNT 70).  The workload runs clean first.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
SHAPE=${SHAPE:-cfg2}
OUT=${OUT:-r02_train_${SHAPE}_kernels}
mkdir -p gpurun_out
SHAPE=$SHAPE python tools/prof_train.py > gpurun_out/p_$SHAPE.log 2>&1 && \
SHAPE=$SHAPE ncu --metrics $M --clock-control none -s ${START:-70} -c ${COUNT:-46} --csv --log-file gpurun_out/$OUT.csv python tools/prof_train.py > gpurun_out/n_$SHAPE.log 2>&1
tail -n 2 gpurun_out/p_$SHAPE.log gpurun_out/n_$SHAPE.log
