# Builds libmstts_b200.so (sm_100a only) in-tree.  `python __graft_entry__.py build` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
CSRC := multi_speaker_tts_b200/csrc
OUT  := multi_speaker_tts_b200/libmstts_b200.so
SRCS := $(wildcard $(CSRC)/*.cu)
OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v \
           --expt-relaxed-constexpr -Iinclude

all: $(OUT)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/mstts_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(OUT): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS)

clean:
	rm -rf build $(OUT)

# standalone hardware probes (run on the GPU box): make probes && build/umma_probe && build/stream_probe
probes: build/umma_probe build/stream_probe build/mma_probe
build/%_probe: tools/%_probe.cu multi_speaker_tts_b200/csrc/sm100_ptx.cuh
	@mkdir -p build
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o $@ $<
