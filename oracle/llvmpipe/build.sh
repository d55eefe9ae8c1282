#!/bin/sh
# build.sh — the display-less Xlib stand-in for Mesa's xlib libGL (llvmpipe) into oracle/_ref/fakex/.  TEST INFRASTRUCTURE ONLY.
# libXext.so.6 is an empty library with the right soname (its three XShm* symbols live in the libX11 stand-in; the dynamic
# linker only needs the DT_NEEDED name to resolve).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_ref/fakex"
mkdir -p "$OUT"
CC=${CC:-gcc}
$CC -O1 -fPIC -shared -Wall -Wextra -Wl,-soname,libX11.so.6 -o "$OUT/libX11.so.6" "$HERE/fakex11.c"
$CC -O1 -fPIC -shared -Wl,-soname,libXext.so.6 -o "$OUT/libXext.so.6" -x c /dev/null
echo "built: $OUT/libX11.so.6 $OUT/libXext.so.6"
