# multi-GPU round: usage bash tools/gpu_multi.sh TAG N   (run under gpurun --gpus N)
tag=$1; n=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_gpus.txt
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
run() { # config steps extra...
  c=$1; k=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $c --steps $k --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_n${n}_cfg$c$3.json 2> gpurun_out/${tag}_n${n}_cfg$c$3.err || echo "cfg $c failed"
}
run 3 100
run 4 64
run 5 20
timeout 300 python bench.py --gpus 1 --config 5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n1_cfg5.json 2> gpurun_out/${tag}_n1_cfg5.err
tail -3 gpurun_out/${tag}_tests.log
python tools/stages.py gpurun_out/${tag}_n*_cfg*.json
