# multi-GPU round (gpurun --gpus N): the multi-GPU parity tests, then the driver's bench command (cfg 3 views + the cfg-5 range-sharded
# leg) and, for comparison, cfg 5 with the round-1 all-reduce merge.  usage: bash tools/gpu_multi.sh TAG N [quick|notests]
tag=$1; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_gpus.txt
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
if [ "$3" != "notests" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
  tail -5 gpurun_out/${tag}_tests.log
fi
if [ "$3" != "quick" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n${n}_default.json 2> gpurun_out/${tag}_n${n}_default.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $n --config 5 --merge allreduce --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n${n}_cfg5_allreduce.json 2> gpurun_out/${tag}_n${n}_cfg5_allreduce.err
  python tools/stages.py gpurun_out/${tag}_n${n}_default.json gpurun_out/${tag}_n${n}_cfg5_allreduce.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_n${n}_default.json").read().strip().splitlines()[-1])
    print(json.dumps(d.get("range_sharded"), indent=1)[:3000])
except Exception as e:
    print("no range leg:", e)
PY
  tail -5 gpurun_out/${tag}_n${n}_default.err
fi
