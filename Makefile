# Builds the product library (CUDA kernels + C-ABI + host C++ mirror) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := $(if $(wildcard /usr/bin/g++),/usr/bin/g++,g++)
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v
LIBDIR := jpgenc_b200/lib
CU := $(wildcard jpgenc_b200/csrc/*.cu)
HOSTSRC := jpgenc_b200/host/huffman_build.cpp jpgenc_b200/host/jfif_writer.cpp jpgenc_b200/host/ppm_reader.cpp \
           jpgenc_b200/host/Image.cpp jpgenc_b200/host/Huffman.cpp
OBJ := $(patsubst jpgenc_b200/csrc/%.cu,build/%.o,$(CU)) $(patsubst jpgenc_b200/host/%.cpp,build/host_%.o,$(HOSTSRC))
HDR := $(wildcard jpgenc_b200/csrc/*.cuh) $(wildcard jpgenc_b200/csrc/*.hpp) $(wildcard jpgenc_b200/host/*.hpp) $(wildcard jpgenc_b200/host/include/*.hpp) include/jpgenc_b200.h

all: $(LIBDIR)/libjpgenc_b200.so jpgenc_b200/bin/jpgEnc

build/%.o: jpgenc_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

build/host_%.o: jpgenc_b200/host/%.cpp $(HDR)
	@mkdir -p build
	$(HOSTCXX) -O2 -std=c++17 -ffp-contract=off -fPIC -Wall -c $< -o $@

$(LIBDIR)/libjpgenc_b200.so: $(OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJ)

# the reference's command line (src/main.cpp) on top of the mirror + GPU library
jpgenc_b200/bin/jpgEnc: jpgenc_b200/host/main.cpp $(LIBDIR)/libjpgenc_b200.so
	@mkdir -p jpgenc_b200/bin
	$(HOSTCXX) -O2 -std=c++17 -ffp-contract=off jpgenc_b200/host/main.cpp -o $@ -L$(LIBDIR) -ljpgenc_b200 -Wl,-rpath,'$$ORIGIN/../lib'

clean:
	rm -rf build $(LIBDIR)/*.so

.PHONY: all clean
