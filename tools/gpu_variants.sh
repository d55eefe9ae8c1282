# kernel-variant exploration: parity tests on the default build, then bench.py per variant.  usage: bash tools/gpu_variants.sh TAG CONFIG name1 name2 ...
# (variants are built by tools/build_variant.sh into vk_gltf_viewer_b200/variants/; "base" = libvkv.so)
tag=$1; cfg=$2; shift; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
for v in "$@"; do
  lib=variants/libvkv_$v.so; [ "$v" = base ] && lib=libvkv.so
  VKV_LIBVKV=$lib timeout 200 python bench.py --config $cfg --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_var_${cfg}_$v.json 2> gpurun_out/${tag}_var_${cfg}_$v.err || echo "$v failed"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cull" -c 6 -o gpurun_out/${tag}_cull python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_cull.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
python tools/stages.py gpurun_out/${tag}_var_${cfg}_*.json
