G=${G:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $T bench.py --gpus $G > gpurun_out/b$G.log 2> gpurun_out/b$G.err; tail -c 800 gpurun_out/b$G.err; wc -c gpurun_out/b$G.log
SHAPE=cfg2 timeout 200 $T tools/dp_breakdown.py > gpurun_out/dpb_cfg2_$G.log 2>&1; tail -25 gpurun_out/dpb_cfg2_$G.log
SHAPE=cfg4 timeout 200 $T tools/dp_breakdown.py > gpurun_out/dpb_cfg4_$G.log 2>&1; tail -25 gpurun_out/dpb_cfg4_$G.log
