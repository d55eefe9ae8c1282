// gltf.cpp — glTF 2.0 binary (GLB) ingest for the host input generators (not part of the per-frame path).
//
// The reference loads assets through fastgltf (assets.cpp:526-552: Options::GenerateMeshIndices | DecomposeNodeMatrices, extensions
// KHR_mesh_quantization / EXT_meshopt_compression / …) and turns every glTF primitive into the buffers of glsl::Primitive
// (assets.cpp:288-373), every scene node into a transform + MeshletDraws (world.cpp:187-293).  fastgltf's parser cannot be built in
// this image (it needs simdjson, fetched at CMake time — SURVEY D7), so this file is the small GLB / JSON reader the survey allows:
// it reads the SAME fields the reference consumes and feeds them to the same host code as the procedural generators
// (vkvh::build_primitive, vkvh_scene_add_node_*), with fastgltf's conversion rules restated:
//   * POSITION through iterateAccessor<vec3>: per component convertComponent<float, T> (tools.hpp:266-289) — float(x), or
//     max(float(x) / float(T max), -1) for normalized integers (KHR_mesh_quantization), honouring byteStride;
//   * indices through copyFromAccessor<uint32_t> (u8 / u16 / u32 widened); a primitive without indices gets 0..n-1
//     (Options::GenerateMeshIndices);
//   * the primitive's AABB from the accessor's min / max (assets.cpp:303-306), zero when absent (getAccessorMinMax's default);
//   * material index + 1, 0 = default material (assets.cpp:292-294, 486-492); baseColorFactor, alphaCutoff, doubleSided;
//   * node matrices decomposed into TRS exactly like fastgltf::math::decomposeTransformMatrix (math.hpp:854-891), nodes walked
//     depth first from scenes[scene].nodes, one transform per node WITH a mesh, all primitives of the mesh sharing it
//     (world.cpp:242-264).
// Not read (nothing on the geometry path consumes them): textures, samplers, animations, skins, cameras, lights, sparse accessors
// (rejected), non-triangle modes (skipped, as the mesh shader path only draws triangle lists).  EXT_meshopt_compression views
// (CompressedBufferDataAdapter, assets.cpp:70-171) are decoded through the caller's decoder (vkvh_set_meshopt_decoder: libvkv's device
// decoder vkv_meshopt_* in this repo); without one they are refused with a message that says so.
#include "scene.hpp"

#include <cstdio>
#include <deque>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <atomic>
#include <thread>
#include <algorithm>

namespace vkvh {
namespace {

// ---- a minimal JSON document ---------------------------------------------------------------------------------------------
struct Json {
	enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
	double num = 0;
	bool b = false;
	std::string str;
	std::vector<Json> arr;
	std::vector<std::pair<std::string, Json>> obj;

	const Json* get(const char* key) const {
		if (kind != Object) return nullptr;
		for (const auto& kv : obj)
			if (kv.first == key) return &kv.second;
		return nullptr;
	}
	bool has(const char* key) const { return get(key) != nullptr; }
	double number(const char* key, double dflt) const { const Json* j = get(key); return (j && j->kind == Number) ? j->num : dflt; }
	long long integer(const char* key, long long dflt) const {
		const Json* j = get(key);
		if (!j || j->kind != Number) return dflt;
		if (!(j->num > -4.0e18)) return (long long)-4000000000000000000LL; // also NaN
		return j->num < 4.0e18 ? (long long)j->num : 4000000000000000000LL;
	}
	bool boolean(const char* key, bool dflt) const { const Json* j = get(key); return (j && j->kind == Bool) ? j->b : dflt; }
	size_t size() const { return kind == Array ? arr.size() : 0; }
};

class JsonParser {
public:
	JsonParser(const char* p, size_t n) : p_(p), end_(p + n) {}
	bool parse(Json& out) { skip(); return value(out, 0) && (skip(), true); }
private:
	void skip() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_; }
	bool lit(const char* s) { size_t n = std::strlen(s); if ((size_t)(end_ - p_) < n || std::memcmp(p_, s, n)) return false; p_ += n; return true; }
	bool string(std::string& out) {
		if (p_ >= end_ || *p_ != '"') return false;
		++p_;
		while (p_ < end_ && *p_ != '"') {
			if (*p_ == '\\') {
				if (++p_ >= end_) return false;
				switch (*p_) {
				case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
				case 'b': out += '\b'; break; case 'f': out += '\f'; break;
				case 'u': { // keep the BMP code point as UTF-8 (names only; nothing numeric depends on it)
					if (end_ - p_ < 5) return false;
					unsigned cp = (unsigned)std::strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
					p_ += 4;
					if (cp < 0x80) out += (char)cp;
					else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
					else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
					break;
				}
				default: out += *p_;
				}
				++p_;
			} else out += *p_++;
		}
		if (p_ >= end_) return false;
		++p_;
		return true;
	}
	bool value(Json& out, int depth) {
		if (depth > 64 || p_ >= end_) return false;
		const char c = *p_;
		if (c == '{') {
			out.kind = Json::Object; ++p_; skip();
			if (p_ < end_ && *p_ == '}') { ++p_; return true; }
			for (;;) {
				std::string key; skip();
				if (!string(key)) return false;
				skip(); if (p_ >= end_ || *p_ != ':') return false; ++p_; skip();
				out.obj.emplace_back(std::move(key), Json());
				if (!value(out.obj.back().second, depth + 1)) return false;
				skip(); if (p_ >= end_) return false;
				if (*p_ == ',') { ++p_; continue; }
				if (*p_ == '}') { ++p_; return true; }
				return false;
			}
		}
		if (c == '[') {
			out.kind = Json::Array; ++p_; skip();
			if (p_ < end_ && *p_ == ']') { ++p_; return true; }
			for (;;) {
				out.arr.emplace_back(); skip();
				if (!value(out.arr.back(), depth + 1)) return false;
				skip(); if (p_ >= end_) return false;
				if (*p_ == ',') { ++p_; continue; }
				if (*p_ == ']') { ++p_; return true; }
				return false;
			}
		}
		if (c == '"') { out.kind = Json::String; return string(out.str); }
		if (c == 't') { out.kind = Json::Bool; out.b = true; return lit("true"); }
		if (c == 'f') { out.kind = Json::Bool; out.b = false; return lit("false"); }
		if (c == 'n') { out.kind = Json::Null; return lit("null"); }
		char* e = nullptr;
		std::string tmp(p_, (size_t)std::min<ptrdiff_t>(end_ - p_, 64));
		out.num = std::strtod(tmp.c_str(), &e);
		if (e == tmp.c_str()) return false;
		out.kind = Json::Number;
		p_ += e - tmp.c_str();
		return true;
	}
	const char* p_; const char* end_;
};

struct Fail { std::string msg; };
[[noreturn]] void fail(const std::string& m) { throw Fail{m}; }

// an array element used as an index (children, scene roots): anything that is not a small non-negative number is "out of range"
size_t index_of(const Json& j) { return (j.kind == Json::Number && j.num >= 0.0 && j.num < 1.0e12) ? (size_t)j.num : (size_t)-1; }

// a byte offset / length / element count from the document: never negative, and small enough (2^40) that the bounds arithmetic below
// (offset + count * stride with stride <= 2^10) cannot wrap
size_t usize(const Json& j, const char* key) {
	const long long v = j.integer(key, 0);
	if (v < 0 || v > (1LL << 40)) fail(std::string(key) + " is negative or beyond 2^40");
	return (size_t)v;
}

std::vector<uint8_t> base64(const std::string& s, size_t from) {
	std::vector<uint8_t> out;
	unsigned acc = 0; int bits = 0;
	for (size_t i = from; i < s.size(); ++i) {
		const char c = s[i];
		int v;
		if (c >= 'A' && c <= 'Z') v = c - 'A'; else if (c >= 'a' && c <= 'z') v = c - 'a' + 26; else if (c >= '0' && c <= '9') v = c - '0' + 52;
		else if (c == '+') v = 62; else if (c == '/') v = 63; else if (c == '=') break; else continue;
		acc = (acc << 6) | (unsigned)v; bits += 6;
		if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
	}
	return out;
}

struct View { const uint8_t* data; size_t size; size_t stride; }; // stride 0 = tightly packed

// EXT_meshopt_compression: the decoder is the caller's (vkvh_set_meshopt_decoder) — in this repo libvkv's device decoder
vkvh_meshopt_decode_fn g_meshopt_decode = nullptr;
void* g_meshopt_decode_user = nullptr;
struct Acc { View view; size_t offset, count; int ctype; bool normalized; int comps; const Json* json; };

size_t ctype_size(int t) { return (t == 5120 || t == 5121) ? 1 : (t == 5122 || t == 5123) ? 2 : (t == 5125 || t == 5126) ? 4 : 0; }

// fastgltf::internal::convertComponent<float, T> (tools.hpp:266-289)
float to_float(const uint8_t* p, int ctype, bool normalized) {
	switch (ctype) {
	case 5120: { int8_t v; std::memcpy(&v, p, 1); return normalized ? std::max((float)v / 127.0f, -1.0f) : (float)v; }
	case 5121: { uint8_t v = *p; return normalized ? (float)v / 255.0f : (float)v; }
	case 5122: { int16_t v; std::memcpy(&v, p, 2); return normalized ? std::max((float)v / 32767.0f, -1.0f) : (float)v; }
	case 5123: { uint16_t v; std::memcpy(&v, p, 2); return normalized ? (float)v / 65535.0f : (float)v; }
	case 5125: { uint32_t v; std::memcpy(&v, p, 4); return normalized ? (float)v / 4294967295.0f : (float)v; }
	default: { float v; std::memcpy(&v, p, 4); return v; }
	}
}

// fastgltf::math::decomposeTransformMatrix (math.hpp:854-891), operation for operation
void decompose(const float* m16, float t[3], float r[4], float s[3]) {
	float m[16];
	std::memcpy(m, m16, 64);
	t[0] = m[12]; t[1] = m[13]; t[2] = m[14];
	m[12] = m[13] = m[14] = 0.f;
	for (int c = 0; c < 3; ++c) {
		const float* col = m + c * 4;
		float sum = col[0] * col[0];
		for (int i = 1; i < 4; ++i) sum += col[i] * col[i];
		s[c] = std::sqrt(sum);
	}
	for (int c = 0; c < 3; ++c)
		for (int i = 0; i < 4; ++i) m[c * 4 + i] /= s[c];
	auto at = [&](int c, int rr) { return m[c * 4 + rr]; };
	float q[4] = {std::max(.0f, 1.f + at(0, 0) - at(1, 1) - at(2, 2)), std::max(.0f, 1.f - at(0, 0) + at(1, 1) - at(2, 2)),
	              std::max(.0f, 1.f - at(0, 0) - at(1, 1) + at(2, 2)), std::max(.0f, 1.f + at(0, 0) + at(1, 1) + at(2, 2))};
	for (int i = 0; i < 4; ++i) q[i] = static_cast<float>(std::sqrt(static_cast<double>(q[i]))) / 2;
	q[0] = std::copysignf(q[0], at(1, 2) - at(2, 1));
	q[1] = std::copysignf(q[1], at(2, 0) - at(0, 2));
	q[2] = std::copysignf(q[2], at(0, 1) - at(1, 0));
	std::memcpy(r, q, 16);
}

// fastgltf::URI::fspath(): the percent-decoded path of a local uri
std::string percent_decode(const std::string& u) {
	std::string out;
	for (size_t i = 0; i < u.size(); ++i) {
		auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; };
		if (u[i] == '%' && i + 2 < u.size() && hex(u[i + 1]) >= 0 && hex(u[i + 2]) >= 0) { out.push_back((char)(hex(u[i + 1]) * 16 + hex(u[i + 2]))); i += 2; }
		else out.push_back(u[i]);
	}
	return out;
}

bool read_file(const std::string& path, std::vector<uint8_t>& out, size_t limit = (size_t)-1) {
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f) return false;
	std::fseek(f, 0, SEEK_END);
	long n = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	if (n < 0) { std::fclose(f); return false; }
	size_t want = (size_t)n < limit ? (size_t)n : limit;
	out.resize(want);
	const size_t got = want ? std::fread(out.data(), 1, want, f) : 0;
	std::fclose(f);
	return got == want;
}

struct Loader {
	const uint8_t* bin = nullptr; size_t binSize = 0;
	Json doc;
	std::deque<std::vector<uint8_t>> owned; // decoded data-URI buffers, external files, densified sparse accessors (addresses stay put)
	std::vector<View> buffers;
	std::map<long long, View> decoded;      // EXT_meshopt_compression views already decoded (several accessors share a view)
	std::string baseDir;                    // folder of the asset file (assetPath.parent_path(), assets.cpp:548,558); empty = no file access
	bool haveDir = false;

	static size_t view_stride(const Json& bv) { // glTF 2.0: 4 .. 252; anything up to 2^10 is taken, more is refused (it would let offsets wrap)
		const size_t st = usize(bv, "byteStride");
		if (st > 1024) fail("bufferView.byteStride beyond 1024");
		return st;
	}
	View bufferView(long long idx) {
		const Json* bvs = doc.get("bufferViews");
		if (!bvs || idx < 0 || (size_t)idx >= bvs->size()) fail("bufferView index out of range");
		const Json& bv = bvs->arr[(size_t)idx];
		if (const Json* ext = bv.get("extensions"))
			if (const Json* mc = ext->get("EXT_meshopt_compression")) {
				// CompressedBufferDataAdapter::ExecuteRange (assets.cpp:111-171): the view's bytes are the decoded stream, count * byteStride of them
				auto hit = decoded.find(idx);
				if (hit != decoded.end()) return hit->second;
				if (!g_meshopt_decode)
					fail("bufferView " + std::to_string(idx) + " is EXT_meshopt_compression-compressed: install a decoder (vkvh_set_meshopt_decoder; libvkv's device decoder "
					     "vkv_meshopt_plan_create / vkv_meshopt_run takes the compressed views as they are)");
				const long long cb = mc->integer("buffer", -1);
				if (cb < 0 || (size_t)cb >= buffers.size() || !buffers[(size_t)cb].data) fail("EXT_meshopt_compression.buffer out of range or without data");
				const size_t coff = usize(*mc, "byteOffset"), clen = usize(*mc, "byteLength");
				if (coff > buffers[(size_t)cb].size || clen > buffers[(size_t)cb].size - coff) fail("EXT_meshopt_compression view exceeds its buffer");
				const size_t stride = usize(*mc, "byteStride"), count = usize(*mc, "count");
				const Json* jm = mc->get("mode"); const Json* jf = mc->get("filter");
				const std::string mode = jm && jm->kind == Json::String ? jm->str : "", filter = jf && jf->kind == Json::String ? jf->str : "NONE";
				const int m = mode == "ATTRIBUTES" ? 0 : mode == "TRIANGLES" ? 1 : mode == "INDICES" ? 2 : -1;
				const int f = filter == "NONE" ? 0 : filter == "OCTAHEDRAL" ? 1 : filter == "QUATERNION" ? 2 : filter == "EXPONENTIAL" ? 3 : -1;
				// the extension's own constraints (and what meshoptimizer asserts): attributes 4-byte multiples up to 256, indices 2 or 4 bytes,
				// whole triangles, filters only on attributes with the strides they are defined for
				if (m < 0 || f < 0) fail("EXT_meshopt_compression with an unknown mode / filter");
				if (m == 0 ? (stride == 0 || stride % 4 != 0 || stride > 256) : (stride != 2 && stride != 4)) fail("EXT_meshopt_compression byteStride not valid for its mode");
				if (m == 1 && count % 3 != 0) fail("EXT_meshopt_compression TRIANGLES count is not a multiple of 3");
				if (f != 0 && m != 0) fail("EXT_meshopt_compression filter on an index view");
				if ((f == 1 && stride != 4 && stride != 8) || (f == 2 && stride != 8)) fail("EXT_meshopt_compression filter with a byteStride it is not defined for");
				// no meshoptimizer stream expands a thousandfold (vertex blocks: >= 1 header bit pair per 16 bytes; index codes: >= 1 byte per triangle):
				// a view that claims more is refused before anything is allocated for it
				if (count > ((size_t)1 << 31) / stride || count * stride > ((clen + 64) << 10)) fail("EXT_meshopt_compression view claims more output than its stream can hold");
				owned.emplace_back(count * stride + 4); // + 4: filters and index decoders may touch whole words at the tail
				const int rc = g_meshopt_decode(g_meshopt_decode_user, (uint32_t)m, (uint32_t)f, (uint32_t)count, (uint32_t)stride, buffers[(size_t)cb].data + coff, clen,
				                                owned.back().data());
				// (the reference ignores the decoders' return codes, assets.cpp:149-150; a stream that does not decode is refused here)
				if (rc != 0) fail("bufferView " + std::to_string(idx) + ": EXT_meshopt_compression stream did not decode (code " + std::to_string(rc) + ")");
				const View v{owned.back().data(), count * stride, view_stride(bv)};
				decoded.emplace(idx, v);
				return v;
			}
		const long long b = bv.integer("buffer", -1);
		if (b < 0 || (size_t)b >= buffers.size()) fail("bufferView.buffer out of range");
		if (!buffers[(size_t)b].data) fail("bufferView reads a fallback buffer (EXT_meshopt_compression placeholder without bytes)");
		const size_t off = usize(bv, "byteOffset"), len = usize(bv, "byteLength");
		if (off > buffers[(size_t)b].size || len > buffers[(size_t)b].size - off) fail("bufferView exceeds its buffer");
		return View{buffers[(size_t)b].data + off, len, view_stride(bv)};
	}
	Acc accessor(long long idx) {
		const Json* as = doc.get("accessors");
		if (!as || idx < 0 || (size_t)idx >= as->size()) fail("accessor index out of range");
		const Json& a = as->arr[(size_t)idx];
		const Json* sparse = a.get("sparse");
		if (!a.has("bufferView") && !sparse) fail("accessor without bufferView");
		Acc r;
		r.json = &a;
		r.offset = usize(a, "byteOffset");
		r.count = usize(a, "count");
		r.ctype = (int)a.integer("componentType", 0);
		r.normalized = a.boolean("normalized", false);
		const Json* ty = a.get("type");
		const std::string t = ty && ty->kind == Json::String ? ty->str : "";
		r.comps = t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : 0;
		const size_t cs = ctype_size(r.ctype);
		if (!cs || !r.comps) fail("accessor with an unsupported componentType / type");
		const size_t elem = cs * r.comps;
		if (a.has("bufferView")) {
			r.view = bufferView(a.integer("bufferView", -1));
			const size_t stride = r.view.stride ? r.view.stride : elem;
			if (r.count && (r.offset > r.view.size || (r.count - 1) * stride + elem > r.view.size - r.offset)) fail("accessor exceeds its bufferView");
		} else r.view = View{nullptr, 0, 0};
		if (sparse) {
			// glTF 2.0 §3.6.2.3 as fastgltf's iterateAccessor reads it (tools.hpp: the sparse index / value pairs override the elements of
			// the base view, or of zeros when the accessor has no bufferView): densified here into a tightly packed copy
			const size_t n = usize(*sparse, "count");
			const Json* si = sparse->get("indices"); const Json* sv = sparse->get("values");
			if (!si || !sv) fail("sparse accessor without indices / values");
			const int ict = (int)si->integer("componentType", 0);
			const size_t is = ctype_size(ict);
			if (ict != 5121 && ict != 5123 && ict != 5125) fail("sparse accessor: indices must be u8 / u16 / u32");
			const View iv = bufferView(si->integer("bufferView", -1)), vv = bufferView(sv->integer("bufferView", -1));
			const size_t io = usize(*si, "byteOffset"), vo = usize(*sv, "byteOffset");
			if (n > r.count || io > iv.size || n * is > iv.size - io || vo > vv.size || n * elem > vv.size - vo) fail("sparse accessor exceeds its bufferViews");
			// (an accessor WITH a bufferView is bounded by the bytes of the file; one made of zeros only is bounded here)
			if (!r.view.data && r.count * elem > ((size_t)1 << 30)) fail("sparse accessor without a bufferView beyond 1 GiB");
			owned.emplace_back(r.count * elem, (uint8_t)0);
			std::vector<uint8_t>& dense = owned.back();
			if (r.view.data) {
				const size_t stride = r.view.stride ? r.view.stride : elem;
				for (size_t i = 0; i < r.count; ++i) std::memcpy(&dense[i * elem], r.view.data + r.offset + i * stride, elem);
			}
			for (size_t k = 0; k < n; ++k) {
				uint32_t at = 0;
				std::memcpy(&at, iv.data + io + k * is, is);
				if (at >= r.count) fail("sparse accessor: index out of range");
				std::memcpy(&dense[(size_t)at * elem], vv.data + vo + k * elem, elem);
			}
			r.view = View{dense.data(), dense.size(), 0};
			r.offset = 0;
		}
		return r;
	}
};

} // namespace
} // namespace vkvh

void vkvh_set_meshopt_decoder(vkvh_meshopt_decode_fn fn, void* user) {
	vkvh::g_meshopt_decode = fn;
	vkvh::g_meshopt_decode_user = user;
}

// everything after the container: buffers (BIN chunk, data URIs, files beside the asset), then materials / meshes / nodes / scene
static vkvh_scene* load_document(vkvh::Loader& L) {
	using namespace vkvh;
	vkvh_scene* s = nullptr;
	try {
		if (const Json* bufs = L.doc.get("buffers"))
			for (size_t i = 0; i < bufs->size(); ++i) {
				const Json& b = bufs->arr[i];
				const Json* uri = b.get("uri");
				const size_t byteLength = usize(b, "byteLength");
				const Json* bext = b.get("extensions");
				const Json* bmc = bext ? bext->get("EXT_meshopt_compression") : nullptr;
				if (!uri && bmc && bmc->boolean("fallback", false)) {
					// fastgltf::sources::Fallback: a placeholder for decoders without the extension; it has no bytes and nothing may read it
					L.buffers.push_back(View{nullptr, byteLength, 0});
				} else if (!uri) { // the GLB-stored buffer
					if (i != 0 || !L.bin) fail("buffer without uri that is not the GLB BIN chunk");
					if (byteLength > L.binSize) fail("buffers[0].byteLength exceeds the BIN chunk");
					L.buffers.push_back(View{L.bin, byteLength, 0});
				} else {
					const std::string& u = uri->str;
					if (u.compare(0, 5, "data:") == 0) {
						const size_t comma = u.find(',');
						if (comma == std::string::npos || u.find(";base64") == std::string::npos) fail("data uri that is not base64");
						L.owned.push_back(base64(u, comma + 1));
					} else {
						// BufferLoadTask (assets.cpp:36-68): folder / uri.fspath(), byteLength bytes; only local files
						if (!L.haveDir) fail("external buffer uri '" + u.substr(0, 40) + "': no asset folder (load the file with vkvh_scene_load_file)");
						if (u.find("://") != std::string::npos) fail("external buffer uri '" + u.substr(0, 40) + "': only local files are read (assets.cpp:52)");
						std::string path = percent_decode(u);
						if (path.empty() || path[0] != '/') path = L.baseDir + "/" + path;
						L.owned.emplace_back();
						if (!read_file(path, L.owned.back(), byteLength) || L.owned.back().size() < byteLength) fail("Failed to open buffer: " + path);
					}
					if (L.owned.back().size() < byteLength) fail("buffer shorter than its byteLength");
					L.buffers.push_back(View{L.owned.back().data(), byteLength ? byteLength : L.owned.back().size(), 0});
				}
			}

		s = vkvh_scene_new();
		// materials (world.cpp:132-178 / assets.cpp:486-492: glTF material i -> index i + 1)
		if (const Json* mats = L.doc.get("materials"))
			for (const Json& m : mats->arr) {
				float albedo[4] = {1, 1, 1, 1};
				if (const Json* pbr = m.get("pbrMetallicRoughness"))
					if (const Json* f = pbr->get("baseColorFactor"))
						for (size_t k = 0; k < 4 && k < f->size(); ++k) albedo[k] = (float)f->arr[k].num;
				const uint32_t idx = vkvh_scene_add_material(s, albedo, m.boolean("doubleSided", false) ? 1 : 0);
				s->materials[idx].alphaCutoff = (float)m.number("alphaCutoff", 0.5);
			}

		// meshes -> primitives (assets.cpp:288-373), flattened in mesh order like the reference's primitiveBuffers.  Two phases, like the
		// reference's PrimitiveProcessingTask (assets.cpp:192-215,375-430: one task-set partition per mesh on the enkiTS pool): the accessors
		// are read here in order (cheap, and every refusal is raised in document order), then the meshlet builds — the load-time
		// dominator, ~1 us per triangle per core — run on all cores, and the results are appended in document order again.
		struct Job {
			size_t mesh;
			std::vector<vkv_Vertex> verts;
			std::vector<uint32_t> idx;
			uint32_t materialIndex;
			float center[3], extents[3];
			std::vector<int16_t> qpos;
			bool qnormalized = false;
			PrimitiveData built;
			bool ok = false, oom = false;
		};
		std::deque<Job> jobs;
		std::vector<std::vector<int32_t>> meshPrims;
		if (const Json* meshes = L.doc.get("meshes"))
			for (const Json& mesh : meshes->arr) {
				meshPrims.emplace_back();
				const Json* prims = mesh.get("primitives");
				if (!prims) continue;
				for (const Json& pr : prims->arr) {
					if (pr.integer("mode", 4) != 4) continue; // only triangle lists reach the mesh-shader path
					const Json* attrs = pr.get("attributes");
					if (!attrs || !attrs->has("POSITION")) fail("primitive without POSITION");
					const Acc pos = L.accessor(attrs->integer("POSITION", -1));
					if (pos.comps != 3) fail("POSITION accessor is not VEC3");
					const size_t cs = ctype_size(pos.ctype), stride = pos.view.stride ? pos.view.stride : cs * 3;
					jobs.emplace_back();
					Job& job = jobs.back();
					job.mesh = meshPrims.size() - 1;
					std::vector<vkv_Vertex>& verts = job.verts;
					verts.resize(pos.count);
					for (size_t i = 0; i < pos.count; ++i) {
						std::memset(&verts[i], 0, sizeof(vkv_Vertex));
						const uint8_t* e = pos.view.data + pos.offset + i * stride;
						for (int k = 0; k < 3; ++k) verts[i].position[k] = to_float(e + k * cs, pos.ctype, pos.normalized);
						verts[i].color[0] = verts[i].color[1] = verts[i].color[2] = verts[i].color[3] = 255;
					}
					std::vector<uint32_t>& idx = job.idx;
					if (pr.has("indices")) {
						const Acc ia = L.accessor(pr.integer("indices", -1));
						if (ia.comps != 1 || (ia.ctype != 5121 && ia.ctype != 5123 && ia.ctype != 5125)) fail("index accessor must be SCALAR u8 / u16 / u32");
						const size_t is = ctype_size(ia.ctype), istride = ia.view.stride ? ia.view.stride : is;
						idx.resize(ia.count);
						for (size_t i = 0; i < ia.count; ++i) {
							const uint8_t* e = ia.view.data + ia.offset + i * istride;
							uint32_t v = 0;
							std::memcpy(&v, e, is);
							idx[i] = v;
						}
					} else { // Options::GenerateMeshIndices
						idx.resize(pos.count);
						for (size_t i = 0; i < pos.count; ++i) idx[i] = (uint32_t)i;
					}
					const long long mat = pr.integer("material", -1);
					job.materialIndex = mat >= 0 ? (uint32_t)mat + 1 : 0u;
					if (job.materialIndex >= s->materials.size()) fail("primitive.material out of range");
					// assets.cpp:303-306: the primitive's AABB comes from the accessor's min / max (zero vectors when absent)
					float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
					const Json* jmin = pos.json->get("min"); const Json* jmax = pos.json->get("max");
					if (jmin && jmin->size() == 3) for (int k = 0; k < 3; ++k) mn[k] = (float)jmin->arr[k].num;
					if (jmax && jmax->size() == 3) for (int k = 0; k < 3; ++k) mx[k] = (float)jmax->arr[k].num;
					for (int k = 0; k < 3; ++k) {
						job.center[k] = (mn[k] + mx[k]) / 2.f;
						job.extents[k] = mx[k] - job.center[k];
					}
					if (pos.ctype == 5122) { // SHORT positions: keep the 16-bit form for the in-register dequantisation (vkv_set_quantized_positions)
						job.qpos.assign(pos.count * 4, 0);
						for (size_t i = 0; i < pos.count; ++i) std::memcpy(&job.qpos[i * 4], pos.view.data + pos.offset + i * stride, 6);
						job.qnormalized = pos.normalized;
					}
				}
			}
		{
			std::atomic<size_t> next{0};
			auto worker = [&]() {
				for (size_t i; (i = next.fetch_add(1)) < jobs.size();) {
					Job& job = jobs[i];
					try { // nothing may unwind out of a worker thread
						job.ok = build_primitive(job.built, std::move(job.verts), job.idx.data(), (uint32_t)job.idx.size(), job.materialIndex);
					} catch (const std::exception&) { job.ok = false; job.oom = true; }
					std::vector<uint32_t>().swap(job.idx);
				}
			};
			// an injected meshlet builder (tests: the reference's meshoptimizer through ctypes) is called from this thread only
			const unsigned nt = builder_is_injected() ? 1u : std::max(1u, std::min({32u, std::thread::hardware_concurrency(), (unsigned)jobs.size()}));
			std::vector<std::thread> pool;
			for (unsigned t = 1; t < nt; ++t) pool.emplace_back(worker);
			worker();
			for (auto& th : pool) th.join();
		}
		for (Job& job : jobs) {
			if (job.oom) fail("out of memory while building a primitive's meshlets");
			if (!job.ok) fail("primitive with no triangles or an index out of range");
			PrimitiveData& pd = job.built;
			for (int k = 0; k < 3; ++k) { pd.header.aabbCenter[k] = job.center[k]; pd.header.aabbExtents[k] = job.extents[k]; }
			pd.qpos = std::move(job.qpos);
			pd.qnormalized = job.qnormalized;
			meshPrims[job.mesh].push_back(add_built_primitive(s, std::move(pd)));
		}
		jobs.clear();

		// nodes (world.cpp:187-228): TRS as given, matrices decomposed (Options::DecomposeNodeMatrices); scene roots in order
		const Json* nodes = L.doc.get("nodes");
		const size_t nNodes = nodes ? nodes->size() : 0;
		std::vector<char> placed(nNodes, 0); // glTF 2.0 §3.5.2: the hierarchy is a strict tree — a node reached twice would make the walk exponential
		std::function<void(size_t, int32_t, int)> addNode = [&](size_t ni, int32_t parent, int depth) {
			if (ni >= nNodes || depth > 256) fail("node index out of range or node hierarchy too deep / cyclic");
			if (placed[ni]) fail("node " + std::to_string(ni) + " has more than one parent (the node hierarchy must be a tree)");
			placed[ni] = 1;
			const Json& n = nodes->arr[ni];
			float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, sc[3] = {1, 1, 1};
			if (const Json* m = n.get("matrix")) {
				if (m->size() != 16) fail("node.matrix must have 16 elements");
				float mm[16];
				for (int k = 0; k < 16; ++k) mm[k] = (float)m->arr[(size_t)k].num;
				decompose(mm, t, r, sc);
			} else {
				if (const Json* v = n.get("translation")) for (size_t k = 0; k < 3 && k < v->size(); ++k) t[k] = (float)v->arr[k].num;
				if (const Json* v = n.get("rotation")) for (size_t k = 0; k < 4 && k < v->size(); ++k) r[k] = (float)v->arr[k].num;
				if (const Json* v = n.get("scale")) for (size_t k = 0; k < 3 && k < v->size(); ++k) sc[k] = (float)v->arr[k].num;
			}
			const long long mesh = n.integer("mesh", -1);
			if (mesh >= 0 && (size_t)mesh >= meshPrims.size()) fail("node.mesh out of range");
			const int32_t self = vkvh_scene_add_node_mesh(s, parent, mesh >= 0 ? meshPrims[(size_t)mesh].data() : nullptr,
			                                             mesh >= 0 ? (uint32_t)meshPrims[(size_t)mesh].size() : 0u, mesh >= 0 ? 1 : 0, t, r, sc);
			if (self < 0) fail("invalid node");
			if (const Json* ch = n.get("children"))
				for (const Json& c : ch->arr) addNode(index_of(c), self, depth + 1);
		};
		const Json* scenes = L.doc.get("scenes");
		if (scenes && scenes->size()) {
			const long long sceneIndex = L.doc.integer("scene", 0);
			if (sceneIndex < 0 || (size_t)sceneIndex >= scenes->size()) fail("scene index out of range");
			const size_t si = (size_t)sceneIndex;
			if (const Json* roots = scenes->arr[si].get("nodes"))
				for (const Json& rn : roots->arr) addNode(index_of(rn), -1, 0);
		}
		if (vkvh_scene_finalize(s) != 0) fail("the draw list exceeds 2^25 MeshletDraws (visbuffer.h.glsl:15-17)");
		return s;
	} catch (...) {
		if (s) vkvh_scene_free(s);
		throw;
	}
}

// GLB container (12-byte header + JSON chunk + optional BIN chunk) or a bare .gltf JSON document -> L.doc / L.bin
static void open_container(vkvh::Loader& L, const void* data, size_t bytes, bool allowJson) {
	using namespace vkvh;
	const uint8_t* p = (const uint8_t*)data;
	auto u32 = [&](size_t o) { uint32_t v; std::memcpy(&v, p + o, 4); return v; };
	const bool glb = p && bytes >= 4 && u32(0) == 0x46546C67u;
	const char* json = nullptr; size_t jsonLen = 0;
	if (!glb) {
		if (!allowJson || !p || !bytes) fail("not a GLB container (magic)");
		json = (const char*)p; jsonLen = bytes;
		if (jsonLen >= 3 && p[0] == 0xEF && p[1] == 0xBB && p[2] == 0xBF) { json += 3; jsonLen -= 3; } // UTF-8 byte order mark
	} else {
		if (bytes < 20) fail("not a GLB container (magic)");
		if (u32(4) != 2) fail("GLB version " + std::to_string(u32(4)) + " (only 2 is supported)");
		const size_t total = u32(8);
		if (total > bytes) fail("GLB length field exceeds the data");
		size_t o = 12;
		while (o + 8 <= total) {
			const size_t len = u32(o), type = u32(o + 4);
			if (o + 8 + len > total) fail("GLB chunk exceeds the container");
			if (type == 0x4E4F534Au && !json) { json = (const char*)p + o + 8; jsonLen = len; }
			else if (type == 0x004E4942u && !L.bin) { L.bin = p + o + 8; L.binSize = len; }
			o += 8 + ((len + 3) & ~(size_t)3);
		}
		if (!json) fail("GLB without a JSON chunk");
	}
	if (!JsonParser(json, jsonLen).parse(L.doc) || L.doc.kind != Json::Object) fail("malformed glTF JSON");
}

extern "C" {

vkvh_scene* vkvh_scene_load_glb(const void* data, size_t bytes, char* err, size_t errcap) {
	auto report = [&](const std::string& m) { if (err && errcap) std::snprintf(err, errcap, "%s", m.c_str()); };
	try {
		vkvh::Loader L;
		open_container(L, data, bytes, false);
		return load_document(L);
	} catch (const vkvh::Fail& f) {
		report(f.msg);
	} catch (const std::exception& e) {
		report(std::string("glTF load failed: ") + e.what());
	}
	return nullptr;
}

// AssetLoadTask::loadGltf (assets.cpp:526-552: MappedGltfFile::FromPath + Parser::loadGltf(file, assetPath.parent_path(), ...), which takes
// .gltf and .glb alike) + BufferLoadTask (assets.cpp:36-68: external buffers read from files beside the asset)
vkvh_scene* vkvh_scene_load_file(const char* path, char* err, size_t errcap) {
	auto report = [&](const std::string& m) { if (err && errcap) std::snprintf(err, errcap, "%s", m.c_str()); };
	try {
		if (!path) vkvh::fail("Failed to open glTF file");
		std::vector<uint8_t> file;
		if (!vkvh::read_file(path, file)) vkvh::fail(std::string("Failed to open glTF file: ") + path);   // assets.cpp:529-531
		vkvh::Loader L;
		const std::string p(path);
		const size_t slash = p.find_last_of('/');
		L.baseDir = slash == std::string::npos ? "." : (slash == 0 ? "/" : p.substr(0, slash));
		L.haveDir = true;
		open_container(L, file.data(), file.size(), true);
		return load_document(L);   // (the scene copies what it keeps: `file` may go)
	} catch (const vkvh::Fail& f) {
		report(f.msg);
	} catch (const std::exception& e) {
		report(std::string("glTF load failed: ") + e.what());
	}
	return nullptr;
}

} // extern "C"
