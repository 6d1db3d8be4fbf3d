# Build libbnr.so (sm_100a only) and the oracle's C pieces.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
PKG := bayesiannetworkregression.jl_b200
CSRC := $(PKG)/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
OBJS := $(CSRC)/bnr_small_kernels.o $(CSRC)/bnr_linalg.o $(CSRC)/bnr_diagnostics.o $(CSRC)/bnr_api.o $(CSRC)/bnr_fit.o
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/bnr.h

all: $(PKG)/libbnr.so

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/libbnr.so: $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart -ldl

clean:
	rm -f $(OBJS) $(PKG)/libbnr.so

.PHONY: all clean
