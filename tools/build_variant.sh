#!/bin/sh
# tools/build_variant.sh NAME -DFOO=1 ... : another in-tree build of libvkv.so with tuning macros -> vk_gltf_viewer_b200/variants/libvkv_NAME.so
# (select with VKV_LIBVKV=variants/libvkv_NAME.so; artefacts are git-ignored but travel to the GPU box)
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v "$@" \
  -shared -o vk_gltf_viewer_b200/variants/libvkv_$name.so vk_gltf_viewer_b200/csrc/*.cu -Iinclude -ldl 2> /tmp/variant_$name.log || { cat /tmp/variant_$name.log; exit 1; }
grep -A2 -E "raster_kernel|cull_kernel" /tmp/variant_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
