# Build the C-ABI library (and the GEMM bring-up harness) for sm_100a. `python -c "import __graft_entry__ as g; g.build()"`
# runs the same commands.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC := thinkdiff_mlre_b200/csrc
LIB := thinkdiff_mlre_b200/libthinkdiff_b200.so

all: $(LIB)

$(LIB): $(CSRC)/td_api.cu $(wildcard $(CSRC)/*.cuh) include/thinkdiff_b200.h
	$(NVCC) $(NVCCFLAGS) -shared -Iinclude -o $@ $(CSRC)/td_api.cu

test_gemm: $(CSRC)/test_gemm.cu $(wildcard $(CSRC)/*.cuh)
	$(NVCC) $(NVCCFLAGS) -o $(CSRC)/test_gemm.bin $(CSRC)/test_gemm.cu

clean:
	rm -f $(LIB) $(CSRC)/test_gemm.bin
