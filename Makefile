# Build everything in-tree: harness (host, g++), oracle (gcc), device library (nvcc, sm_100a).
PKG := 3d-dycoreplanet_b200
NVCC := /usr/local/cuda/bin/nvcc
CXX := g++
CXXFLAGS := -O3 -march=x86-64-v3 -std=c++17 -fopenmp -fPIC -Wall -Wno-unused-variable
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-fopenmp -Xptxas -v
CU_SRCS := $(wildcard $(PKG)/csrc/device/*.cu)
CU_HDRS := $(wildcard $(PKG)/csrc/device/*.cuh) $(wildcard $(PKG)/csrc/device/*.h) include/dcp.h

all: lib/libdcp_harness.so oracle lib/libdcp.so tests/cpp/host_mirror_test tests/cpp/halo_test

lib/libdcp_harness.so: $(wildcard $(PKG)/csrc/harness/*.hpp) $(PKG)/csrc/harness/harness_api.cpp include/dcp_harness.h
	mkdir -p lib
	$(CXX) $(CXXFLAGS) -shared -o $@ $(PKG)/csrc/harness/harness_api.cpp

oracle:
	$(MAKE) -C oracle

CU_OBJS := $(patsubst $(PKG)/csrc/device/%.cu,build/obj/%.o,$(CU_SRCS))

build/obj/%.o: $(PKG)/csrc/device/%.cu $(CU_HDRS)
	mkdir -p build/obj build/ptxas
	$(NVCC) $(NVFLAGS) -c -o $@ $< > build/ptxas/$*.log 2>&1 || (cat build/ptxas/$*.log; false)

lib/libdcp.so: $(CU_OBJS)
	mkdir -p lib
	$(NVCC) -shared -Xcompiler -fopenmp -o $@ $(CU_OBJS) -lcudart -lgomp -ldl
	cat build/ptxas/*.log > build/ptxas.log

# C++ host mirror (include/dcp.hpp) checked against the oracle; the oracle is linked as the checker only
tests/cpp/host_mirror_test: tests/cpp/host_mirror_test.cpp include/dcp.hpp include/dcp.h include/dcp_harness.h lib/libdcp.so lib/libdcp_harness.so oracle
	$(CXX) -std=c++17 -O2 -Wall -Wextra -I include -o $@ tests/cpp/host_mirror_test.cpp -L lib -ldcp -ldcp_harness -L oracle/_build -loracle \
	  -Wl,-rpath,'$$ORIGIN/../../lib' -Wl,-rpath,'$$ORIGIN/../../oracle/_build'

# two ranks through the multi-GPU part of the ABI (needs two GPUs to run)
tests/cpp/halo_test: tests/cpp/halo_test.cpp include/dcp.h lib/libdcp.so
	$(CXX) -std=c++17 -O2 -Wall -Wextra -I include -o $@ tests/cpp/halo_test.cpp -L lib -ldcp -Wl,-rpath,'$$ORIGIN/../../lib'

clean:
	rm -rf lib build oracle/_build tests/cpp/host_mirror_test tests/cpp/halo_test
.PHONY: all oracle clean
