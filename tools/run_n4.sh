for G in 4 2; do
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G"
timeout 400 $T bench.py --gpus $G > gpurun_out/b$G.log 2> gpurun_out/b$G.err; tail -c 300 gpurun_out/b$G.err; wc -c gpurun_out/b$G.log
done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
timeout 200 $T bench.py --gpus 4 --impl reference --steps 2 > gpurun_out/ref4.log 2> gpurun_out/ref4.err; cut -c1-400 gpurun_out/ref4.log
