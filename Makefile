# Builds libonmf_b200.so (sm_100a only) in-tree, and the oracle's C restatement.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
CSRC      := onmf_ontf_ndl_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB       := onmf_ontf_ndl_b200/libonmf_b200.so

all: $(LIB) oracle

build/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh include/onmf_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

build/lars.o: $(CSRC)/lars_fast.cuh

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

oracle: oracle/liblars_oracle.so

oracle/liblars_oracle.so: oracle/lars_oracle.c
	gcc -O2 -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIB) oracle/liblars_oracle.so

.PHONY: all oracle clean variant

# experimental variant of the coder: make variant NAME=pf DEFS="-DLARS_PREFETCH=1" -> build/variants/libonmf_b200_pf.so
variant: $(OBJS)
	@mkdir -p build/$(NAME) build/variants
	$(NVCC) $(NVFLAGS) $(DEFS) -c $(CSRC)/lars.cu -o build/$(NAME)/lars.o 2> build/$(NAME)/lars.ptxas.log || (cat build/$(NAME)/lars.ptxas.log; exit 1)
	$(NVCC) $(ARCH) -shared -o build/variants/libonmf_b200_$(NAME).so $(filter-out build/lars.o,$(OBJS)) build/$(NAME)/lars.o -lcudart
