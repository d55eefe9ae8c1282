# one GPU round: parity tests, benches of cfg 1-4 (cfg 3 with both meshlet builders), launch list, ncu full capture.
# usage: bash tools/gpu_round.sh TAG [notests] [noncu]
tag=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt
if [ "$2" != "notests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
fi
timeout 400 python bench.py > gpurun_out/${tag}_bench3.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --meshlets morton --no-cpu-baseline > gpurun_out/${tag}_bench3_morton.json 2>> gpurun_out/${tag}_bench.err
for c in 1 2 4; do
  timeout 300 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench$c.json 2>> gpurun_out/${tag}_bench.err
done
if [ "$3" != "noncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_b.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|cull_kernel|hiz" -c 12 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
fi
tail -3 gpurun_out/${tag}_tests.log; tail -1 gpurun_out/${tag}_smoke.log
python tools/stages.py gpurun_out/${tag}_bench3.json gpurun_out/${tag}_bench3_morton.json gpurun_out/${tag}_bench[124].json
tail -5 gpurun_out/${tag}_bench.err
