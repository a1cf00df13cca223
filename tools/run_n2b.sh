T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/check_dp_training.py > gpurun_out/dp_check2.log 2>&1; tail -6 gpurun_out/dp_check2.log | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "parallel or dp or DP" > gpurun_out/t2.log 2>&1; tail -3 gpurun_out/t2.log
timeout 400 $T bench.py --gpus 2 --steps 20 --blocks train,cfg4 > gpurun_out/b2b.log 2> gpurun_out/b2b.err; tail -c 600 gpurun_out/b2b.err; wc -c gpurun_out/b2b.log
B200REC_PEER_ALLREDUCE=0 timeout 400 $T bench.py --gpus 2 --steps 20 --blocks train > gpurun_out/b2c.log 2> gpurun_out/b2c.err; wc -c gpurun_out/b2c.log
