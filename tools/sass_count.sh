#!/bin/sh
# static SASS instruction count per kernel of a library: tools/sass_count.sh vk_gltf_viewer_b200/libvkv.so
cuobjdump -sass "$1" 2>/dev/null | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -n
