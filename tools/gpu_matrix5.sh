# cfg 5 on one GPU (1.02 B triangles, 8K), variants side by side.  usage: bash tools/gpu_matrix5.sh TAG name1 name2 ...
tag=$1; shift
for v in "$@"; do
  lib=variants/libvkv_$v.so; [ "$v" = base ] && lib=libvkv.so
  VKV_LIBVKV=$lib timeout 300 python bench.py --config 5 --shard views --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_m5_$v.json 2> gpurun_out/${tag}_m5_$v.err || echo "$v failed"
done
python tools/stages.py gpurun_out/${tag}_m5_*.json
