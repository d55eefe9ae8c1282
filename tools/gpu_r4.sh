# round-2 late validation job (one GPU): parity tests, smoke, the default bench line (frames in flight in e2e), cfg 1/2/4, the reference arm,
# raster knob A/Bs (tools/build_variant.sh), launch list + ncu full capture of the final kernels.  usage: bash tools/gpu_r4.sh TAG [variants...]
tag=${1:-r4a}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/${tag}_bench3.json 2> gpurun_out/${tag}_bench.err
for c in 1 2 4; do
  timeout 300 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench$c.json 2>> gpurun_out/${tag}_bench.err
done
timeout 300 python bench.py --config 3 --steps 64 --no-cpu-baseline --cone-cull > gpurun_out/${tag}_bench3_cone.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref3.json 2>> gpurun_out/${tag}_bench.err
# knob A/Bs: batch* on cfg 3, serial* on cfg 4 (same job, same box as the base lines above / below)
for v in "$@"; do
  case $v in batch*) c=3;; *) c=4;; esac
  VKV_LIBVKV=variants/libvkv_$v.so timeout 200 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_var_${c}_$v.json 2>> gpurun_out/${tag}_bench.err || echo "$v failed"
done
timeout 300 python bench.py --config 3 --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_var_3_base.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|raster_big|cull_kernel|hiz" -c 16 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_tests.log; tail -1 gpurun_out/${tag}_smoke.log
python tools/stages.py gpurun_out/${tag}_bench3.json gpurun_out/${tag}_bench[124].json gpurun_out/${tag}_bench3_cone.json gpurun_out/${tag}_var_*.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench3.json").read().strip().splitlines()[-1])
print("e2e", d["e2e"]["value"], "one-in-flight", d["e2e"].get("one_frame_in_flight"), "readback", d["e2e_readback"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"))
print("roofline", json.dumps(d["roofline"])[:400])
PY
tail -5 gpurun_out/${tag}_bench.err
