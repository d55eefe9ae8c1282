# multi-GPU round (gpurun --gpus N): the multi-GPU parity tests, then range-sharded cfg-5 and view-sharded cfg-3 benches.  usage: bash tools/gpu_multi.sh TAG N [quick]
tag=$1; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_gpus.txt
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
if [ "$3" != "quick" ]; then
  port=29561
  for cfg in 5 3; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $cfg --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n${n}_cfg$cfg.json 2> gpurun_out/${tag}_n${n}_cfg$cfg.err
    port=$((port+1))
  done
  python tools/stages.py gpurun_out/${tag}_n${n}_cfg5.json gpurun_out/${tag}_n${n}_cfg3.json
fi
