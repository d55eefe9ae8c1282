# usage: bash tools/gpu_variants.sh "<config>" name1 name2 ...   (name "base" = libvkv.so)
cfg=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=variants/libvkv_$v.so; [ "$v" = base ] && lib=libvkv.so
  VKV_LIBVKV=$lib python bench.py --config $cfg --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/var_${cfg}_$v.json 2> gpurun_out/var_${cfg}_$v.err || echo "$v failed"
done
python tools/stages.py gpurun_out/var_${cfg}_*.json
