# Multi-GPU run of one box: `gpurun --gpus $G -- 'G=2 WHAT="tests dp sharded bench breakdown ref" bash tools/run_multi.sh'`
# (the command lines behind profiles/r02_bench_n{2,4,8}.json, r02_dp_exactness_n2.txt, r02_dp_breakdown_cfg*_n8.txt and
# r02_reference_arm_n4.json).  WHAT selects the steps; BLOCKS / STEPS are handed to bench.py.
G=${G:-2}
WHAT=${WHAT:-"dp sharded bench"}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G"
mkdir -p gpurun_out
for w in $WHAT; do
  case $w in
    tests)     timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t$G.log 2>&1; tail -3 gpurun_out/t$G.log ;;
    dp)        timeout 300 $T tools/check_dp_training.py > gpurun_out/dp_check$G.log 2>&1; tail -12 gpurun_out/dp_check$G.log | cut -c1-400 ;;
    sharded)   timeout 300 $T tools/check_sharded.py > gpurun_out/sh_check$G.log 2>&1; tail -4 gpurun_out/sh_check$G.log ;;
    bench)     timeout 600 $T bench.py --gpus $G ${STEPS:+--steps $STEPS} ${BLOCKS:+--blocks $BLOCKS} > gpurun_out/b$G.log 2> gpurun_out/b$G.err
               tail -c 800 gpurun_out/b$G.err; wc -c gpurun_out/b$G.log ;;
    breakdown) for S in cfg2 cfg4; do SHAPE=$S timeout 200 $T tools/dp_breakdown.py > gpurun_out/dpb_${S}_$G.log 2>&1; tail -25 gpurun_out/dpb_${S}_$G.log; done ;;
    ref)       timeout 300 $T bench.py --gpus $G --impl reference --steps 3 > gpurun_out/ref$G.log 2> gpurun_out/ref$G.err; cut -c1-400 gpurun_out/ref$G.log ;;
  esac
done
