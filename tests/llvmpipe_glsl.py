"""Reference GLSL, executed as GLSL: the arithmetic of the reference's task and mesh shaders, taken verbatim from /root/reference/shaders at
test time (never copied into the repo), compiled by Mesa's GLSL compiler and run by llvmpipe as a fragment shader over one texel per
MeshletDraw / triangle (tests/llvmpipe_lib.py::compute).  Only what is NOT arithmetic is replaced: buffer-reference fetches become texelFetch
of textures the test fills, the built-in outputs become colour outputs.  TEST INFRASTRUCTURE ONLY.

Rewrites of the reference text (mechanical, arithmetic untouched):
  * `for (uint i = 0; i < N;` -> `for (uint i = 0u; i < Nu;` in culling.h.glsl: GLSL 1.40 (this Mesa's compat profile) has no implicit int -> uint
    conversion (GLSL 4.00 added it);
  * the `#include "common.h.glsl"` line is replaced by that file's own GLSL branch (the macros PARAMETER_COPY, GLSL_NAMESPACE_BEGIN, ...).
GL_ARB_shading_language_420pack supplies the `const` locals with non-constant initialisers the reference uses throughout.
"""
import os
import re

import numpy as np

REF = "/root/reference/shaders"


def available():
    return os.path.isdir(REF)


def ref_lines(path, first, last):
    return "".join(open(os.path.join(REF, path)).read().splitlines(keepends=True)[first - 1:last])


def prelude():
    common = open(os.path.join(REF, "common.h.glsl")).read()
    glsl_branch = common[common.index("#else") + 5:common.rindex("#endif", 0, common.rindex("#endif"))]
    return "#version 140\n#extension GL_ARB_shading_language_420pack : require\n" + glsl_branch


def culling_header():
    text = open(os.path.join(REF, "culling.h.glsl")).read().replace('#include "common.h.glsl"', "")
    out = re.sub(r"for \(uint i = 0; i < (\d+);", r"for (uint i = 0u; i < \1u;", text)
    assert out.count("0u; i <") == 2, "culling.h.glsl: the loop rewrite did not apply"
    return out


def task_shader():
    """culling.h.glsl whole + visbuffer.task.glsl:50-52 (world box, frustum test) and :56-61 (projected box, mip level, sample position)"""
    return prelude() + culling_header() + f"""
struct MeshletIn {{ vec3 aabbCenter; vec3 aabbExtents; }};
struct CameraIn {{ vec4 frustum[6]; mat4 prevOcclusionViewProjection; }};
uniform CameraIn camera;
uniform ivec2 pyramidSize;
uniform int mode;
uniform sampler2D xf0, xf1, xf2, xf3, boxC, boxE;
out vec4 color;
void main() {{
	ivec2 at = ivec2(gl_FragCoord.xy);
	mat4 transformMatrix = mat4(texelFetch(xf0, at, 0), texelFetch(xf1, at, 0), texelFetch(xf2, at, 0), texelFetch(xf3, at, 0));
	MeshletIn meshlet = MeshletIn(texelFetch(boxC, at, 0).xyz, texelFetch(boxE, at, 0).xyz);
{ref_lines("visbuffer/visbuffer.task.glsl", 50, 52)}
	float levelOut = 0.0, zmaxOut = 0.0; vec2 centerOut = vec2(0.0);
	if (visible) {{
{ref_lines("visbuffer/visbuffer.task.glsl", 56, 61)}
		levelOut = level; centerOut = projectedCenter; zmaxOut = projectedAabb[1].z;
	}}
	color = mode == 0 ? vec4(visible ? 1.0 : 0.0, levelOut, centerOut) : vec4(zmaxOut, worldAabbCenter);
}}
"""


def mesh_shader():
    """visbuffer.mesh.glsl:44 (mvp), :61 (clip position), :71 (transformDet), :90-98 (the facing decision), one triangle per texel"""
    facing = ref_lines("visbuffer/visbuffer.mesh.glsl", 90, 98)
    facing = facing.replace("clipVertices[indices.x]", "clipVertices[0]").replace("clipVertices[indices.y]", "clipVertices[1]").replace("clipVertices[indices.z]", "clipVertices[2]")
    facing = facing.replace("gl_MeshPrimitivesEXT[pidx].gl_CullPrimitiveEXT =", "culled =")
    assert facing.count("culled =") == 2 and "indices" not in facing
    vertex = ref_lines("visbuffer/visbuffer.mesh.glsl", 61, 61)
    return prelude() + f"""
struct CameraIn {{ mat4 viewProjection; }};
struct PushIn {{ CameraIn camera; }};
struct CameraRef {{ PushIn cameraBuffer; }};
uniform mat4 viewProjection;
uniform int mode;
uniform sampler2D xf0, xf1, xf2, xf3, p0, p1, p2;
out vec4 color;
struct VertexIn {{ vec3 position; }};
void main() {{
	ivec2 at = ivec2(gl_FragCoord.xy);
	mat4 transformMatrix = mat4(texelFetch(xf0, at, 0), texelFetch(xf1, at, 0), texelFetch(xf2, at, 0), texelFetch(xf3, at, 0));
	CameraRef pushConstants; pushConstants.cameraBuffer.camera.viewProjection = viewProjection;
{ref_lines("visbuffer/visbuffer.mesh.glsl", 44, 44)}
	vec3 clipVertices[3]; vec4 clip[3];
	for (int k = 0; k < 3; ++k) {{
		VertexIn vertex = VertexIn(k == 0 ? texelFetch(p0, at, 0).xyz : k == 1 ? texelFetch(p1, at, 0).xyz : texelFetch(p2, at, 0).xyz);
{vertex}
		clip[k] = pos; clipVertices[k] = pos.xyw;
	}}
{ref_lines("visbuffer/visbuffer.mesh.glsl", 71, 71)}
	bool culled; float detOut;
	{{
{facing}
		detOut = det;
	}}
	color = mode == 0 ? clip[0] : mode == 1 ? clip[1] : mode == 2 ? clip[2] : vec4(detOut, transformDet, culled ? 1.0 : 0.0, 0.0);
}}
"""


def hiz_coordinates_shader():
    """the sample coordinate of hiz_reduce.comp.glsl:28, `(vec2(pos) + vec2(0.5)) / pushConstants.imageSize`, cut out of the reference's line and
    evaluated per texel of a target of the level's size (pos = gl_GlobalInvocationID.xy there, the fragment's integer position here)"""
    line = ref_lines("hiz_reduce.comp.glsl", 28, 28)
    m = re.search(r"texture\(sampled_textures_heap\[pushConstants\.sourceImage\], (.*)\)\.x;", line)
    assert m, "hiz_reduce.comp.glsl:28 is not the line this test was written against"
    return prelude() + f"""
struct HiZReducePushConstants {{ uvec2 imageSize; }};
uniform HiZReducePushConstants pushConstants;
out vec4 color;
void main() {{
	uvec2 pos = uvec2(gl_FragCoord.xy);
	color = vec4({m.group(1)}, 0.0, 1.0);
}}
"""


def to_tex(a, Wc):
    """[n, k<=4] -> ([Hc, Wc, 4] float32, Hc)"""
    n = a.shape[0]
    Hc = max(1, -(-n // Wc))
    t = np.zeros((Hc * Wc, 4), np.float32)
    t[:n, :a.shape[1]] = a
    return t.reshape(Hc, Wc, 4)
