# Build without Python (the same commands kitti_motion_compensation_b200/build.py runs).
#   make            libkmc_b200.so (CUDA kernels + C ABI, sm_100a) and libkitti_motion_compensation_lib.so (C++ mirror)
#   make example    lib/motion_compensate_runs (the reference's CLI on top of the mirror)
#   make oracle     oracle/libkmc_oracle.so (CPU restatement) and, where /root/reference exists, oracle/_ref/libkmc_ref.so (the
#                   reference's own sources compiled unmodified) — test infrastructure only
#   make cpp-tests  tests/cpp/_build/test_{dropin_host,dropin_gpu,eigen_shim}
#   make ref-tests  the reference's own test/*.cpp, unmodified, against the drop-in (tests/cpp/_ref_build/build/)
NVCC      ?= nvcc
CXX       ?= g++
PKG       := kitti_motion_compensation_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/lib
INC       := include
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC,-fvisibility=hidden -cudart static
CXXFLAGS  := -std=c++17 -O2 -fPIC -Wall -Wextra

KERNEL_SRC := $(CSRC)/kmc_kernels.cu $(CSRC)/kmc_kernels_bulk.cu $(CSRC)/kmc_capi.cu $(CSRC)/kmc_pipeline.cu $(CSRC)/kmc_host_math.cpp $(CSRC)/kmc_run.cpp
KERNEL_HDR := $(CSRC)/kmc_kernels.cuh $(CSRC)/kmc_point_math.cuh $(CSRC)/kmc_host_math.hpp $(CSRC)/kmc_internal.hpp $(CSRC)/kmc_host_pool.hpp $(INC)/kmc_b200.h
MIRROR_HDR := $(wildcard $(INC)/kitti_motion_compensation/*.hpp)

all: $(LIB)/libkmc_b200.so $(LIB)/libkitti_motion_compensation_lib.so

$(LIB)/libkmc_b200.so: $(KERNEL_SRC) $(KERNEL_HDR)
	@mkdir -p $(LIB)
	$(NVCC) $(NVCCFLAGS) -I $(INC) -I $(CSRC) -o $@ $(KERNEL_SRC)

$(LIB)/libkitti_motion_compensation_lib.so: $(CSRC)/kmc_dropin.cpp $(MIRROR_HDR) $(INC)/kmc_b200.h $(LIB)/libkmc_b200.so
	$(CXX) $(CXXFLAGS) -shared -I $(INC) -o $@ $< -L $(LIB) -lkmc_b200 '-Wl,-rpath,$$ORIGIN'

example: $(LIB)/motion_compensate_runs $(LIB)/bench_motion_compensate_frame
$(LIB)/bench_motion_compensate_frame: examples/bench_motion_compensate_frame.cpp $(LIB)/libkitti_motion_compensation_lib.so
	$(CXX) -std=c++17 -O2 -Wall -I $(INC) -o $@ $< -L $(LIB) -lkitti_motion_compensation_lib -lkmc_b200 -lpthread '-Wl,-rpath,$$ORIGIN'
$(LIB)/motion_compensate_runs: examples/motion_compensate_runs.cpp $(LIB)/libkitti_motion_compensation_lib.so
	$(CXX) -std=c++17 -O2 -Wall -I $(INC) -o $@ $< -L $(LIB) -lkitti_motion_compensation_lib -lkmc_b200 '-Wl,-rpath,$$ORIGIN'

oracle:
	$(MAKE) -C oracle
	$(MAKE) -C oracle ref

cpp-tests: all
	@mkdir -p tests/cpp/_build
	for t in test_dropin_host test_dropin_gpu; do \
	  $(CXX) -std=c++17 -O1 -Wall -I $(INC) -I tests/cpp -o tests/cpp/_build/$$t tests/cpp/$$t.cpp -L $(LIB) -lkitti_motion_compensation_lib -lkmc_b200 -lpthread -Wl,-rpath,$(abspath $(LIB)); done
	$(CXX) -std=c++17 -O1 -Wall -Wextra -pedantic -Werror -I $(INC) -I tests/cpp -o tests/cpp/_build/test_eigen_shim tests/cpp/test_eigen_shim.cpp

# the reference's OWN gtest files, unmodified, against the drop-in (needs the reference checkout; binaries in tests/cpp/_ref_build/build,
# run from there so that "../testing_assets" resolves; same as kitti_motion_compensation_b200.build.build_reference_tests())
REF ?= /root/reference
ref-tests: all
	@mkdir -p tests/cpp/_ref_build/build
	@[ -d tests/cpp/_ref_build/testing_assets ] || (mkdir -p tests/cpp/_ref_build && cp -r $(REF)/testing_assets tests/cpp/_ref_build/ && rm -rf tests/cpp/_ref_build/testing_assets/*/*/image_0*)
	for t in test_motion_compensation test_timestamp_mocking test_lie_algebra test_trajectory_interpolation test_oxts_to_pose; do \
	  $(CXX) -std=c++17 -O1 -I tests/cpp/gtest_stub -I $(INC) -o tests/cpp/_ref_build/build/$$t $(REF)/test/$$t.cpp -L $(LIB) -lkitti_motion_compensation_lib -lkmc_b200 -lpthread '-Wl,-rpath,$$ORIGIN/../../../../$(LIB)'; done

clean:
	rm -f $(LIB)/*.so $(LIB)/motion_compensate_runs; rm -rf tests/cpp/_build; $(MAKE) -C oracle clean

.PHONY: all example oracle cpp-tests ref-tests clean
