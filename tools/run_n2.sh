T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; tail -3 gpurun_out/t2.log
timeout 300 $T tools/check_dp_training.py > gpurun_out/dp_check2.log 2>&1; tail -12 gpurun_out/dp_check2.log
timeout 300 $T tools/check_sharded.py > gpurun_out/sh_check2.log 2>&1; tail -4 gpurun_out/sh_check2.log
timeout 600 $T bench.py --gpus 2 > gpurun_out/b2.log 2> gpurun_out/b2.err; tail -c 1500 gpurun_out/b2.err; wc -c gpurun_out/b2.log
