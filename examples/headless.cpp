// headless.cpp — the reference's frame loop against libvkv, display-free, in the reference's own language (C++).
//
// What Application::run() does around the geometry hot path (application.cpp:620-1010), with the Vulkan task / mesh pipeline and the
// hiz_reduce dispatches replaced by the C ABI of include/vkv.h and the asset side by include/vkv_host.h:
//   load the asset (AssetLoadTask, assets.cpp:526-600)            vkvh_scene_load_file  (or a procedural BASELINE scene);
//     EXT_meshopt_compression views (assets.cpp:111-171)            vkvh_set_meshopt_decoder -> vkv_meshopt_* on the device
//   upload buffers, build the draw list (World::addAsset, ...)      vkvh_scene_upload -> vkv_upload
//   per frame: camera + transforms, "Visbuffer pass", "HiZ reduction", frameOverlap frames in flight
//                                                                    vkv_update_staged + vkv_frame_submit / vkv_frame_wait
//   visbuffer resolve (application.cpp:917-949)                     vkv_resolve + vkv_read_color -> a PPM file
// There is no CPU fallback: without a CUDA device vkv_create fails and the program says so (exit code 3).
//
//   usage: headless [--asset file.gltf|file.glb | --scene icosphere|atrium|lattice|city] [--size WxH] [--frames N] [--one-pass] [--out image.ppm]
#include <vkv.h>
#include <vkv_host.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h> // pinned staging memory only (cudaHostAlloc); every GPU operation goes through vkv.h

namespace {
constexpr uint32_t frameOverlap = 3; // application.hpp:146

int upload_cb(void* user, const void* host, size_t bytes, uint64_t* dev) { return vkv_upload(static_cast<vkv_ctx*>(user), host, bytes, dev); }

// EXT_meshopt_compression: what CompressedBufferDataAdapter does on the CPU (assets.cpp:111-171), done by libvkv's device decoder behind the
// host reader's decoder hook — one call per compressed bufferView.  The context is created on first use and then renders the frames.
struct DeviceDecoder { vkv_ctx** vkv; uint32_t W, H; };
int decode_cb(void* user, uint32_t mode, uint32_t filter, uint32_t count, uint32_t stride, const void* src, size_t src_bytes, void* dst) {
	auto* d = static_cast<DeviceDecoder*>(user);
	if (!*d->vkv && vkv_create(d->vkv, 0, d->W, d->H) != VKV_OK) return -100;
	vkv_ctx* c = *d->vkv;
	std::vector<unsigned char> stream(src_bytes + 32, 0); // a little slack behind the stream, as the decoder's tests provide
	std::memcpy(stream.data(), src, src_bytes);
	const size_t bytes = ((size_t)count * stride + 15) & ~(size_t)15;
	const vkv_MeshoptView view = {mode, filter, count, stride, 0, src_bytes, 0};
	uint64_t srcDev = 0, dstDev = 0;
	vkv_meshopt_plan* plan = nullptr;
	int32_t rc = -100;
	if (vkv_upload(c, stream.data(), stream.size(), &srcDev) == VKV_OK && vkv_alloc(c, bytes, &dstDev) == VKV_OK &&
	    vkv_meshopt_plan_create(c, &view, 1, &plan) == VKV_OK && vkv_meshopt_run(c, plan, srcDev, stream.size(), dstDev, bytes) == VKV_OK &&
	    vkv_meshopt_results(c, plan, &rc) == VKV_OK && rc == 0 && vkv_download(c, dstDev, dst, (size_t)count * stride) != VKV_OK)
		rc = -100;
	if (plan) vkv_meshopt_plan_destroy(c, plan);
	if (dstDev) vkv_free(c, dstDev);
	if (srcDev) vkv_free(c, srcDev);
	return rc;
}

[[noreturn]] void die(int code, const char* what, const char* why) {
	std::fprintf(stderr, "headless: %s: %s\n", what, why ? why : "");
	std::exit(code);
}
} // namespace

int main(int argc, char** argv) {
	std::string asset, kind = "icosphere", out;
	uint32_t W = 1280, H = 720, frames = 64;
	uint32_t flags = VKV_FRAME_TWO_PASS;
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto next = [&]() -> const char* { if (i + 1 >= argc) die(2, a.c_str(), "needs a value"); return argv[++i]; };
		if (a == "--asset") asset = next();
		else if (a == "--scene") kind = next();
		else if (a == "--size") { if (std::sscanf(next(), "%ux%u", &W, &H) != 2) die(2, "--size", "expects WxH"); }
		else if (a == "--frames") frames = (uint32_t)std::atoi(next());
		else if (a == "--one-pass") flags = VKV_FRAME_ONE_PASS;
		else if (a == "--out") out = next();
		else die(2, a.c_str(), "unknown argument");
	}

	// ---- the asset side (host only)
	char err[512] = "";
	vkvh_scene* scene = nullptr;
	vkv_ctx* vkv = nullptr;
	DeviceDecoder decoder = {&vkv, W, H};
	vkvh_set_meshopt_decoder(decode_cb, &decoder);
	if (!asset.empty()) scene = vkvh_scene_load_file(asset.c_str(), err, sizeof(err));
	else if (kind == "icosphere") scene = vkvh_scene_icosphere(57);
	else if (kind == "atrium") scene = vkvh_scene_atrium(128);
	else if (kind == "lattice") scene = vkvh_scene_lattice(10, 10, 10, 224, 0x5EED0003ull);
	else if (kind == "city") scene = vkvh_scene_city(50, 40, 10000, 0x5EED0004ull);
	if (!scene) die(2, asset.empty() ? kind.c_str() : asset.c_str(), err[0] ? err : "unknown scene");
	vkvh_counts cnt;
	vkvh_scene_counts(scene, &cnt);
	std::printf("scene: %u meshlet draws, %llu triangles (%llu unique), %u nodes with a transform slot\n", cnt.draws,
	            (unsigned long long)cnt.triangles_instanced, (unsigned long long)cnt.triangles_unique, cnt.transforms);

	// ---- context + uploads
	if (!vkv && vkv_create(&vkv, 0, W, H) != VKV_OK) die(3, "vkv_create", vkv_last_error(nullptr));
	float eye[3], center[3];
	const float up[3] = {0.f, 1.f, 0.f};
	vkv_Camera camera;
	vkvh_scene_default_view(scene, 0, frames, eye, center);
	vkvh_camera_update(&camera, eye, center, up, W, H, 1);
	vkv_VisbufferPushConstants pc;
	if (vkvh_scene_upload(scene, upload_cb, vkv, &camera, &pc) != 0) die(1, "vkvh_scene_upload", vkv_last_error(vkv));

	// one camera / transform buffer and one pinned staging block per frame slot (camera.cpp:86, world.cpp:5-6)
	const size_t transformBytes = (size_t)cnt.transforms * 64;
	uint64_t cameraBuffer[frameOverlap], transformBuffer[frameOverlap];
	vkv_Camera* pinnedCamera[frameOverlap];
	float* pinnedTransforms[frameOverlap];
	for (uint32_t s = 0; s < frameOverlap; ++s) {
		if (vkv_upload(vkv, &camera, sizeof(camera), &cameraBuffer[s]) != VKV_OK || vkv_upload(vkv, vkvh_scene_transforms(scene), transformBytes, &transformBuffer[s]) != VKV_OK)
			die(1, "vkv_upload", vkv_last_error(vkv));
		if (cudaHostAlloc((void**)&pinnedCamera[s], sizeof(vkv_Camera), cudaHostAllocDefault) != cudaSuccess ||
		    cudaHostAlloc((void**)&pinnedTransforms[s], transformBytes ? transformBytes : 64, cudaHostAllocDefault) != cudaSuccess)
			die(1, "cudaHostAlloc", "pinned staging memory");
		if (transformBytes) std::memcpy(pinnedTransforms[s], vkvh_scene_transforms(scene), transformBytes);
	}

	// ---- the frame loop (application.cpp:620-1010 around the hot path)
	uint32_t ticket[frameOverlap] = {};
	uint64_t visible = 0, drained = 0;
	auto retire = [&](uint32_t slot) {
		vkv_stats st;
		if (vkv_frame_wait(vkv, ticket[slot], &st) != VKV_OK) die(1, "vkv_frame_wait", vkv_last_error(vkv));
		ticket[slot] = 0;
		visible += st.visible_a + st.visible_b;
		drained += st.drain_items_a + st.drain_items_b;
	};
	for (uint32_t f = 0; f < frames; ++f) {
		const uint32_t slot = f % frameOverlap;                                  // application.cpp:642
		if (ticket[slot]) retire(slot);                                          // the slot's fence
		vkvh_scene_default_view(scene, f, frames, eye, center);
		vkvh_camera_update(&camera, eye, center, up, W, H, f == 0);              // camera.cpp:170-193
		*pinnedCamera[slot] = camera;
		if (vkv_update_staged(vkv, cameraBuffer[slot], pinnedCamera[slot], sizeof(vkv_Camera)) != VKV_OK ||
		    (transformBytes && vkv_update_staged(vkv, transformBuffer[slot], pinnedTransforms[slot], transformBytes) != VKV_OK))
			die(1, "vkv_update_staged", vkv_last_error(vkv));
		pc.cameraBuffer = cameraBuffer[slot];
		pc.transformBuffer = transformBuffer[slot];
		if (vkv_frame_submit(vkv, &pc, flags, &ticket[slot]) != VKV_OK) die(1, "vkv_frame_submit", vkv_last_error(vkv));
	}
	for (uint32_t k = 0; k < frameOverlap; ++k) {                               // oldest outstanding slot first
		const uint32_t slot = (frames + k) % frameOverlap;
		if (ticket[slot]) retire(slot);
	}
	std::printf("%u frames: %.1f meshlets rasterised per frame, %.1f queued for the drain kernel per frame\n", frames,
	            frames ? (double)visible / frames : 0.0, frames ? (double)drained / frames : 0.0);

	// ---- visbuffer resolve of the last frame (application.cpp:917-949) -> PPM
	if (!out.empty()) {
		if (vkv_resolve(vkv, &pc) != VKV_OK) die(1, "vkv_resolve", vkv_last_error(vkv));
		std::vector<uint32_t> rgba((size_t)W * H);
		if (vkv_read_color(vkv, rgba.data()) != VKV_OK) die(1, "vkv_read_color", vkv_last_error(vkv));
		FILE* fp = std::fopen(out.c_str(), "wb");
		if (!fp) die(1, out.c_str(), "cannot write");
		std::fprintf(fp, "P6\n%u %u\n255\n", W, H);
		for (uint32_t px : rgba) { const unsigned char c[3] = {(unsigned char)(px & 255), (unsigned char)((px >> 8) & 255), (unsigned char)((px >> 16) & 255)}; std::fwrite(c, 1, 3, fp); }
		std::fclose(fp);
		std::printf("wrote %s\n", out.c_str());
	}
	for (uint32_t s = 0; s < frameOverlap; ++s) { cudaFreeHost(pinnedCamera[s]); cudaFreeHost(pinnedTransforms[s]); }
	vkv_destroy(vkv);
	vkvh_scene_free(scene);
	return 0;
}
