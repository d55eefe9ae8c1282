# one GPU round: parity tests, benches of cfg 1-4 (cfg 3 with both meshlet builders), the extension A/Bs (cone cull, int16 positions),
# launch list, ncu full capture.  usage: bash tools/gpu_round.sh TAG [notests] [noncu]
tag=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
fi
timeout 400 python bench.py > gpurun_out/${tag}_bench3.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --meshlets morton --no-cpu-baseline > gpurun_out/${tag}_bench3_morton.json 2>> gpurun_out/${tag}_bench.err
for c in 1 2 4 5; do
  timeout 300 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench$c.json 2>> gpurun_out/${tag}_bench.err
done
# extensions, A/B on the configurations they are meant for
timeout 300 python bench.py --config 1 --steps 64 --no-cpu-baseline --cone-cull > gpurun_out/${tag}_bench1_cone.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config 4 --steps 64 --no-cpu-baseline --cone-cull > gpurun_out/${tag}_bench4_cone.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config 41 --steps 64 --no-cpu-baseline --positions f32 > gpurun_out/${tag}_bench41_f32.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config 41 --steps 64 --no-cpu-baseline --positions i16 > gpurun_out/${tag}_bench41_i16.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config 2 --steps 64 --no-cpu-baseline --positions i16 > gpurun_out/${tag}_bench2_i16.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref3.json 2>> gpurun_out/${tag}_bench.err
if [ "$3" != "noncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_b.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|raster_big|cull_kernel|hiz" -c 16 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
  for pos in f32 i16; do
    timeout 600 ncu --set full --clock-control none -k regex:"raster_kernel" -c 4 -o gpurun_out/${tag}_full41_$pos python bench.py --config 41 --positions $pos --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu41_$pos.log 2>&1
  done
fi
tail -3 gpurun_out/${tag}_tests.log; tail -1 gpurun_out/${tag}_smoke.log
python tools/stages.py gpurun_out/${tag}_bench3.json gpurun_out/${tag}_bench3_morton.json gpurun_out/${tag}_bench[1245].json gpurun_out/${tag}_bench1_cone.json gpurun_out/${tag}_bench4_cone.json gpurun_out/${tag}_bench41_f32.json gpurun_out/${tag}_bench41_i16.json gpurun_out/${tag}_bench2_i16.json
tail -5 gpurun_out/${tag}_bench.err
