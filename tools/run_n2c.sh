T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/check_dp_training.py > gpurun_out/dp_check2.log 2>&1; tail -6 gpurun_out/dp_check2.log | cut -c1-330
timeout 400 $T bench.py --gpus 2 --steps 20 --blocks train,cfg4 > gpurun_out/b2d.log 2> gpurun_out/b2d.err; tail -c 300 gpurun_out/b2d.err; wc -c gpurun_out/b2d.log
