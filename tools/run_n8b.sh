G=${G:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $T bench.py --gpus $G > gpurun_out/b$G.log 2> gpurun_out/b$G.err; tail -c 600 gpurun_out/b$G.err; wc -c gpurun_out/b$G.log
