# Top-level build: libvr.so (CUDA, sm_100a only), the C++ host shim, and the CPU oracle (test infrastructure).
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  := $(shell [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
PKG  := cl_volume_renderer_b200
CSRC := $(PKG)/csrc
ARCH := -gencode arch=compute_100a,code=sm_100a
# --fmad=false: the numeric contract — no contraction, so ray positions match the oracle bit for bit
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC -ccbin $(CXX) -Xptxas -v
CU_SRCS := $(CSRC)/vr_api.cu $(CSRC)/vr_volume_ops.cu $(CSRC)/vr_sdf.cu $(CSRC)/vr_render.cu $(CSRC)/vr_frame_filter.cu $(CSRC)/vr_quiet.cu $(CSRC)/vr_comm.cu
CU_OBJS := $(CU_SRCS:.cu=.o)
# A/B build (tools/ab/libvr_ab.so): the same sources with -DVR_AB — the SDF schedules tried on the way (tools/ab/vr_sdf_variants.cu,
# VR_SDF_MODE), the tile / grid knobs of the wave, the register-budget variants of k_trace_pt.  Not part of the product library.
AB_OBJS := $(patsubst $(CSRC)/%.cu,tools/ab/%.ab.o,$(CU_SRCS)) tools/ab/vr_sdf_variants.ab.o tools/ab/vr_tf_parse.ab.o

all: $(PKG)/libvr.so host oracle ab

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/vr_internal.h $(CSRC)/vr_device.cuh $(CSRC)/vr_sdf_common.cuh include/vr.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

$(CSRC)/vr_tf_parse.o: $(CSRC)/vr_tf_parse.cpp $(CSRC)/vr_internal.h include/vr.h
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

$(PKG)/libvr.so: $(CU_OBJS) $(CSRC)/vr_tf_parse.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -ldl

ab: tools/ab/libvr_ab.so
tools/ab/%.ab.o: $(CSRC)/%.cu $(CSRC)/vr_internal.h $(CSRC)/vr_device.cuh $(CSRC)/vr_sdf_common.cuh include/vr.h
	$(NVCC) $(NVFLAGS) -DVR_AB -I$(CSRC) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)
tools/ab/vr_sdf_variants.ab.o: tools/ab/vr_sdf_variants.cu $(CSRC)/vr_internal.h $(CSRC)/vr_device.cuh $(CSRC)/vr_sdf_common.cuh include/vr.h
	$(NVCC) $(NVFLAGS) -DVR_AB -I$(CSRC) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)
tools/ab/vr_tf_parse.ab.o: $(CSRC)/vr_tf_parse.cpp $(CSRC)/vr_internal.h include/vr.h
	$(NVCC) $(NVFLAGS) -DVR_AB -x cu -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)
tools/ab/libvr_ab.so: $(AB_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -ldl

host: $(PKG)/libvr.so
	@if [ -f $(PKG)/host/Makefile ]; then $(MAKE) -C $(PKG)/host; fi

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.log $(PKG)/libvr.so tools/ab/*.o tools/ab/*.log tools/ab/libvr_ab.so
	$(MAKE) -C oracle clean
.PHONY: all host oracle clean
