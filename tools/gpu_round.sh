mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log
python bench.py --steps 100 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
python bench.py --config 1 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.json 2>> gpurun_out/bench3.err
python bench.py --config 2 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2>> gpurun_out/bench3.err
python bench.py --config 4 --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2>> gpurun_out/bench3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster|cull" -c 16 -o gpurun_out/r1e_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/tests.log; cat gpurun_out/bench3.json | head -c 3000
