# 8-GPU round: strip-mode parity tests (world 4), the driver's bench command at N = 8 (cfg 3 views + cfg-5 range-sharded leg), cfg 5 with the
# round-1 all-reduce merge for comparison.  usage: bash tools/gpu_multi8.sh TAG N
tag=$1; n=${2:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "strips or p2p" > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n${n}_default.json 2> gpurun_out/${tag}_n${n}_default.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $n --config 5 --merge allreduce --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_n${n}_cfg5_allreduce.json 2> gpurun_out/${tag}_n${n}_cfg5_allreduce.err
python tools/stages.py gpurun_out/${tag}_n${n}_default.json gpurun_out/${tag}_n${n}_cfg5_allreduce.json
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_n${n}_default.json").read().strip().splitlines()[-1])
    r = d.get("range_sharded") or {}
    print({k: r.get(k) for k in ("ms_per_frame", "n1_ms_per_frame", "speedup_vs_n1", "merge_parity", "merge_a_ms", "merge_b_ms", "stages_ms_max_over_ranks", "error")})
    print(r.get("nvlink"))
except Exception as e:
    print("no range leg:", e)
PY
tail -3 gpurun_out/${tag}_n${n}_default.err
