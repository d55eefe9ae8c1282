# bench matrix: variants x configs (device-timed frame + stage table).  usage: bash tools/gpu_matrix.sh TAG "cfgs" name1 name2 ...
tag=$1; cfgs=$2; shift; shift
mkdir -p gpurun_out
for c in $cfgs; do
  for v in "$@"; do
    lib=variants/libvkv_$v.so; [ "$v" = base ] && lib=libvkv.so
    VKV_LIBVKV=$lib timeout 200 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_m${c}_$v.json 2> gpurun_out/${tag}_m${c}_$v.err || echo "$v cfg $c failed"
  done
  python tools/stages.py gpurun_out/${tag}_m${c}_*.json
done
