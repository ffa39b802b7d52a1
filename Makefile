# Top-level build: libfxg.so (CUDA, sm_100a only), the drop-in host tools (C) and the test oracle.
NVCC     ?= /usr/local/cuda/bin/nvcc
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ifastx_toolkit_b200/csrc
CSRC     := fastx_toolkit_b200/csrc
LIB      := fastx_toolkit_b200/libfxg.so
CU       := $(CSRC)/fxg_kernels.cu $(CSRC)/fxg_stats.cu $(CSRC)/fxg_stats4.cu $(CSRC)/fxg_clip.cu $(CSRC)/fxg_collapse.cu $(CSRC)/fxg_text.cu $(CSRC)/fxg_deflate.cu $(CSRC)/fxg_extra.cu $(CSRC)/fxg_barcode.cu $(CSRC)/fxg_pipeline.cu $(CSRC)/fxg_comm.cu $(CSRC)/fxg_dcollapse.cu $(CSRC)/fxg_api.cu
HDR      := include/fxg.h include/fxg_synth.h $(CSRC)/fxg_device.cuh $(CSRC)/fxg_kernels.cuh $(CSRC)/fxg_clip_dpx.cuh $(CSRC)/fxg_collapse.cuh $(CSRC)/fxg_comm.h

.PHONY: all lib tools oracle clean ptxas
all: lib tools oracle

OBJDIR   := build/obj
OBJ      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU))

# one object per .cu, compiled side by side (the CUB scans of fxg_pipeline.cu / fxg_collapse.cu dominate).  The library
# depends on the SOURCES, not on the objects: a tree that ships libfxg.so without build/ (the GPU box) is up to date.
lib: $(LIB)
$(LIB): $(CU) $(HDR)
	@$(MAKE) --no-print-directory -j$(shell nproc) objs
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl
.PHONY: objs
objs: $(OBJ)
$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

ptxas: $(CU) $(HDR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $(CSRC)/fxg_kernels.cu -o /dev/null

tools: lib
	@if [ -f $(CSRC)/host/Makefile ]; then $(MAKE) --no-print-directory -C $(CSRC)/host; fi

oracle:
	$(MAKE) --no-print-directory -C oracle all

clean:
	rm -f $(LIB); rm -rf bin build; $(MAKE) -C oracle clean
