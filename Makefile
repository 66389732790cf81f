# Build of the product library (sm_100a only) and of the checkers under oracle/.
NVCC   ?= nvcc
PKG    := cuda-surf_b200
CSRC   := $(PKG)/csrc
LIB    := $(PKG)/libsurfb200.so
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -I$(CSRC)
CU_SRCS := $(CSRC)/integral.cu $(CSRC)/hessian.cu $(CSRC)/nms.cu $(CSRC)/describe.cu $(CSRC)/describe_tma.cu $(CSRC)/match.cu $(CSRC)/postmatch.cu
OBJS    := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS)) build/ctx.o build/synth.o

all: $(LIB) demo

build/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh include/surfb200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
build/ctx.o: $(CSRC)/ctx.cpp $(CSRC)/common.cuh include/surfb200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@
build/synth.o: $(CSRC)/synth.cpp include/surfb200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

# the reference's main.cpp flow on this library (include/compat/surf.h), no OpenCV
demo: $(LIB) examples/surf_demo.cpp include/compat/surf.h
	$(NVCC) -O2 -std=c++17 -Iinclude -Iinclude/compat -o build/surf_demo examples/surf_demo.cpp -L$(PKG) -lsurfb200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

oracle:
	env -u CC $(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)
.PHONY: all oracle clean demo
