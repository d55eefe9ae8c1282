#!/bin/sh
# build_ref.sh — compile the parts of the REFERENCE that are plain C++ into oracle/_ref/ (git-ignored, travels with gpurun).
# TEST INFRASTRUCTURE ONLY.  Sources are compiled where they lie under /root/reference; nothing is copied into the repo.
# Outputs:
#   oracle/_ref/libmeshopt_ref.so  meshoptimizer @ the reference's pinned submodule (meshopt_buildMeshlets & co, the codecs)
#   oracle/_ref/libmeshopt_ref_nosimd.so  the same, -DMESHOPTIMIZER_NO_SIMD (scalar decode filters)
#   oracle/_ref/libref_shim.so     culling.h.glsl (isAabbInFrustum, getWorldSpaceAabbExtent, aabbPositions, projectAabb) and
#                                  visbuffer.task.glsl:57-61 (mip selection), visbuffer.mesh.glsl:44,61,65,71,90-98 (vertex transform,
#                                  determinants, facing decision) + the shared layout headers compiled as C++ against
#                                  the reference's glm; hiz_reduce.comp.glsl:11-15,21-31 + application.cpp:472-473,965,979 (the
#                                  reduce shader's main and its dispatch sizes); camera.cpp:38-48,70-84 (reverseDepth, generateCameraFrustum); glm
#                                  perspective/lookAt; fastgltf::math translate/rotate/scale.
# The GLSL task/mesh/fragment/compute shaders themselves cannot be built or run here (no glslang, no Vulkan ICD) — see DESIGN.md.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
TMP=$(mktemp -d /tmp/vkv_ref_build.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT" "$TMP/vulkan"
CXX=${CXX:-g++}

# 1. meshoptimizer, straight from the submodule
$CXX -O2 -fPIC -shared -o "$OUT/libmeshopt_ref.so" "$REF"/submodules/meshoptimizer/src/*.cpp
# the same sources with the portable scalar filter / codec paths (vertexfilter.cpp:75-160, vertexcodec.cpp:300-415): the definition
# oracle/meshopt_decode.cpp restates (the SSE filter paths round differently and use rsqrtps, whose bits are CPU-vendor specific)
$CXX -O2 -fPIC -shared -ffp-contract=off -DMESHOPTIMIZER_NO_SIMD -o "$OUT/libmeshopt_ref_nosimd.so" "$REF"/submodules/meshoptimizer/src/*.cpp

# 2. shim.  common.h.glsl pulls <vulkan/vk.hpp> (volk, fmt, tracy) only for VkDeviceAddress: give it a one-line stand-in.
printf '#pragma once\n#include <cstdint>\ntypedef std::uint64_t VkDeviceAddress;\n' > "$TMP/vulkan/vk.hpp"
# culling.h.glsl is dual GLSL/C++ up to line 30; the rest (array constructors) is GLSL-only. Compile the C++-valid head in place.
# The only edit is mechanical: GLSL swizzle `plane.xyz` -> `vec3(plane)` (glm has no .xyz member without MS extensions).
{ sed -n '1,30p' "$REF/shaders/culling.h.glsl" | sed 's/plane\.xyz/vec3(plane)/g'; printf 'GLSL_NAMESPACE_END\n#endif\n'; } > "$TMP/culling_head.h.glsl"
# culling.h.glsl:31-56 (aabbPositions, projectAabb) is GLSL-only syntax; four mechanical rewrites make it C++ against glm, the
# arithmetic untouched: the array constructor `vec3[8](...)` -> a braced initialiser, the array return type `vec3[2]` ->
# std::array<vec3, 2>, the swizzle `clip.xy` -> vec2(clip), `return vec3[2](a, b)` -> `return {a, b}`.
{ printf '#pragma once\nGLSL_NAMESPACE_BEGIN\n'
  sed -n '31,56p' "$REF/shaders/culling.h.glsl" | sed \
    -e 's/const vec3 aabbPositions\[8\] = vec3\[8\](/const vec3 aabbPositions[8] = {/' \
    -e 's/^);$/};/' \
    -e 's/^vec3\[2\] projectAabb/inline std::array<vec3, 2> projectAabb/' \
    -e 's/clip\.xy/vec2(clip)/g' \
    -e 's/return vec3\[2\](ssMin, ssMax);/return {ssMin, ssMax};/'
  printf 'GLSL_NAMESPACE_END\n'; } > "$TMP/culling_tail.h.glsl"
# visbuffer.task.glsl:57-61 (mip selection + sample position) as a function body; two rewrites: `X.xy` -> vec2(X), and max( -> glm::max(
# (inside namespace glsl both std::max and glm::max are visible on the C++ side; glm::max(x, y) = x < y ? y : x is GLSL's definition)
{ printf 'GLSL_NAMESPACE_BEGIN\ninline void taskMipAndCenter(const std::array<vec3, 2>& projectedAabb, ivec2 pyramidSize, float& levelOut, vec2& centerOut) {\n'
  sed -n '57,61p' "$REF/shaders/visbuffer/visbuffer.task.glsl" | sed -e 's/\(projectedAabb\[[01]\]\)\.xy/vec2(\1)/g' -e 's/\bmax(/glm::max(/g'
  printf '\tlevelOut = level; centerOut = projectedCenter;\n}\nGLSL_NAMESPACE_END\n'; } > "$TMP/task_lines.inc"
# visbuffer.mesh.glsl: the arithmetic lines of the mesh shader as three inline functions — :44 (mvp), :61 + :65 (clip position, clipVertices),
# :71 + :90-98 (determinants and the facing decision).  Rewrites, arithmetic untouched: the buffer-reference chain
# `pushConstants.cameraBuffer.camera.` -> `camera.`, the swizzle `pos.xyw` -> vec3(pos.x, pos.y, pos.w), `clipVertices[vidx] =` -> `clipVertex =`,
# `clipVertices[indices.k]` -> `cv[indices.k]`, and the built-in output `gl_MeshPrimitivesEXT[pidx].gl_CullPrimitiveEXT =` -> `culled =`.
MESH="$REF/shaders/visbuffer/visbuffer.mesh.glsl"
{ printf 'GLSL_NAMESPACE_BEGIN\ninline mat4 meshMvp(const Camera& camera, const mat4& transformMatrix) {\n'
  sed -n '44p' "$MESH" | sed -e 's/pushConstants\.cameraBuffer\.camera\./camera./'
  printf '\treturn mvp;\n}\ninline vec4 meshVertex(const mat4& mvp, const Vertex& vertex, vec3& clipVertex) {\n'
  sed -n '61p;65p' "$MESH" | sed -e 's/clipVertices\[vidx\] = pos\.xyw;/clipVertex = vec3(pos.x, pos.y, pos.w);/'
  printf '\treturn pos;\n}\ninline float meshTransformDet(const mat4& transformMatrix) {\n'
  sed -n '71p' "$MESH"
  printf '\treturn transformDet;\n}\ninline bool meshCull(const vec3* cv, uvec3 indices, float transformDet, float& detOut) {\n\tbool culled;\n'
  sed -n '90,98p' "$MESH" | sed -e 's/clipVertices\[/cv[/g' -e 's/gl_MeshPrimitivesEXT\[pidx\]\.gl_CullPrimitiveEXT =/culled =/'
  printf '\tdetOut = det;\n\treturn culled;\n}\nGLSL_NAMESPACE_END\n'; } > "$TMP/mesh_lines.inc"
# hiz_reduce.comp.glsl:11-15 (the push-constant struct) and :21-31 (main) against stand-ins for the built-ins (ref_shim.cpp: the two descriptor
# heaps, gl_GlobalInvocationID, texture(), imageStore()).  Three rewrites, arithmetic untouched: `void main()` -> a named inline function,
# the swizzle `gl_GlobalInvocationID.xy` -> uvec2(gl_GlobalInvocationID), and GLSL's implicit uvec2 -> vec2 conversion of the divisor written
# out (`/ pushConstants.imageSize` -> `/ vec2(pushConstants.imageSize)`; glm's operators do not mix element types).
HIZ="$REF/shaders/hiz_reduce.comp.glsl"
{ printf 'GLSL_NAMESPACE_BEGIN\n'
  sed -n '11,15p' "$HIZ"
  printf 'static thread_local HiZReducePushConstants pushConstants;\n'
  sed -n '21,31p' "$HIZ" | sed -e 's/^void main()/inline void hizReduceMain()/' -e 's/gl_GlobalInvocationID\.xy/uvec2(gl_GlobalInvocationID)/' \
      -e 's#/ pushConstants\.imageSize)#/ vec2(pushConstants.imageSize))#'
  printf 'GLSL_NAMESPACE_END\n'; } > "$TMP/hiz_lines.inc"
grep -q 'vec2(pushConstants.imageSize)' "$TMP/hiz_lines.inc" && grep -q 'hizReduceMain' "$TMP/hiz_lines.inc" || { echo "hiz_reduce.comp.glsl: rewrite did not apply" >&2; exit 1; }
# application.cpp:472-473 (mip count of the pyramid), :965 (size of the level dispatch i writes) and the two group-count expressions of
# the vkCmdDispatch at :979, as two inline functions
APP="$REF/src/vk_gltf_viewer/application.cpp"
{ printf 'inline std::uint32_t hizMipLevels(glm::u32vec2 renderResolution) {\n'
  sed -n '472,473p' "$APP"
  printf '\treturn mipLevels;\n}\ninline void hizDispatch(glm::u32vec2 renderResolution, std::uint32_t i, glm::u32vec2& levelSizeOut, glm::u32vec2& groupsOut) {\n'
  sed -n '965p' "$APP"
  sed -n '979p' "$APP" | sed -e 's/vkCmdDispatch(cmd, \(.*\), \(.*\), 1);/groupsOut = glm::u32vec2(\1, \2);/'
  printf '\tlevelSizeOut = levelSize;\n}\n'; } > "$TMP/hiz_dispatch.inc"
grep -q 'groupsOut = glm::u32vec2(fg::alignUp' "$TMP/hiz_dispatch.inc" || { echo "application.cpp:979: rewrite did not apply" >&2; exit 1; }
# srgb.h.glsl is dual GLSL / C++ except for its swizzles: `X.rgb` -> vec3(X) (glm has no .rgb member without swizzle extensions)
sed -e 's/\([A-Za-z]*\)\.rgb/vec3(\1)/g' "$REF/shaders/srgb.h.glsl" > "$TMP/srgb_lines.h.glsl"
# camera.cpp free functions reverseDepth (38-48) and generateCameraFrustum (70-84)
{ printf '#define ZoneScoped\n#include <array>\n#include <glm/glm.hpp>\n'; sed -n '38,48p;70,84p' "$REF/src/vk_gltf_viewer/camera.cpp"; } > "$TMP/camera_fns.inc"

$CXX -std=c++20 -O2 -fPIC -shared -ffp-contract=off -pthread \
    -I"$TMP" -I"$REF/shaders" -I"$REF/submodules/glm" -I"$REF/submodules/fastgltf/include" \
    -o "$OUT/libref_shim.so" "$HERE/ref_shim.cpp"
echo "built: $(ls "$OUT")"
