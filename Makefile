# Build everything in-tree (artefacts are git-ignored but travel to the GPU box with gpurun).
#   make            -> libvkv.so (CUDA, sm_100a) + libvkv_host.so (host generators) + oracle/liboracle.so
#   make examples   -> examples/headless (the reference's frame loop against libvkv, no display)
#   make ref        -> oracle/_ref/*.so from the reference's own sources (only where /root/reference exists) + oracle/_ref/fakex (llvmpipe pin)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
PKG       := vk_gltf_viewer_b200
CSRC      := $(PKG)/csrc
HOSTSRC   := $(PKG)/host

# -fmad=false: the parity contract is "every fp32 op individually rounded" (SURVEY §8c); -prec-div/-prec-sqrt are defaults.
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
             -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -Wextra -ffp-contract=off -fno-fast-math -pthread

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_HDRS   := $(wildcard $(CSRC)/*.cuh) $(wildcard include/*.h)
HOST_SRCS := $(wildcard $(HOSTSRC)/*.cpp)
HOST_HDRS := $(wildcard $(HOSTSRC)/*.hpp) $(wildcard include/*.h)

all: $(PKG)/libvkv.so $(PKG)/libvkv_host.so oracle/liboracle.so

# one object per .cu (so `make -j` compiles them side by side and a change to one kernel rebuilds one file); objects and the
# per-file ptxas logs live in build/ (git-ignored)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS))

build/%.o: $(CSRC)/%.cu $(CU_HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) $(EXTRA_NVCCFLAGS) -c -o $@ $< -Iinclude 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(PKG)/libvkv.so: $(CU_OBJS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CU_OBJS) -ldl
	@cat build/*.ptxas.log > $(PKG)/ptxas.log
	@grep -E "registers|spill" $(PKG)/ptxas.log | sort | uniq -c | sort -rn | head -5 || true

$(PKG)/libvkv_host.so: $(HOST_SRCS) $(HOST_HDRS)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRCS) -Iinclude

# -mavx2, not -march=native: the library is built here and travels to the GPU box, whose host CPU is a different model — an
# instruction set both have, instead of a SIGILL there; -ffp-contract=off stays (it is the parity source)
oracle/liboracle.so: oracle/oracle.cpp oracle/meshopt_decode.cpp oracle/meshlet_build.cpp oracle/accessors.cpp oracle/oracle.h include/vkv_abi.h
	$(CXX) $(CXXFLAGS) -O3 -mavx2 -shared -o $@ oracle/oracle.cpp oracle/meshopt_decode.cpp oracle/meshlet_build.cpp oracle/accessors.cpp

# the reference's frame loop against the C ABI, display-free (C++, the reference's language); links cudart for pinned staging memory only
examples: examples/headless
# (links against the built libraries without listing them as prerequisites: on a box that received prebuilt .so files but no build/
# objects this must not trigger an nvcc rebuild — run `make` first on a fresh checkout)
examples/headless: examples/headless.cpp include/vkv.h include/vkv_host.h
	$(CXX) -O2 -std=c++17 -Wall -Wextra -o $@ $< -Iinclude -I/usr/local/cuda/include -L$(PKG) -lvkv -lvkv_host \
	    -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/../$(PKG)'

ref:
	@if [ -d /root/reference ]; then sh oracle/build_ref.sh; else echo "no /root/reference here: using prebuilt oracle/_ref if present"; fi
	@sh oracle/llvmpipe/build.sh

# the Xlib stand-in that lets tests drive llvmpipe (Mesa's libGL inside Nsight Compute) without an X server; test infrastructure
llvmpipe:
	sh oracle/llvmpipe/build.sh

clean:
	rm -rf $(PKG)/libvkv.so $(PKG)/libvkv_host.so oracle/liboracle.so $(PKG)/ptxas.log build examples/headless

.PHONY: all ref clean examples llvmpipe
