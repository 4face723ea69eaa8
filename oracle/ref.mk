# Builds oracle/_ref/libtbrm_ref.so: the few host-side source files of the REFERENCE that can be compiled outside Unreal Engine,
# taken from where they lie under $(REF) (never copied into this repository), against the engine-type shim in oracle/ue_shim/
# plus our C wrapper (ref_wrap.cpp). Test infrastructure: used by tests/test_ref_pin_cpu.py and tests/golden/make_golden_ref.py
# to pin the oracle's host parameter math and the volume normalisation (row f3) to the reference's
# own code. The shaders (HLSL) and the RHI drivers cannot be built here; see DESIGN.md §2.
#     make -f ref.mk            (run from oracle/; build() of __graft_entry__.py does it when /root/reference exists)
REF ?= /root/reference
CXX := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
CXXFLAGS := -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -w
INC := -I ue_shim -I $(REF)/Source/Raymarcher/Public -I $(REF)/Source/VolumeTextureToolkit/Public
SRC := $(REF)/Source/Raymarcher/Private/Rendering/LightingShaderUtils.cpp \
       $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/VolumeInfo.cpp
_ref/libtbrm_ref.so: ref_wrap.cpp $(SRC) $(wildcard ue_shim/*.h) tbrm_oracle.h ../include/tbrm.h
	mkdir -p _ref
	$(CXX) $(CXXFLAGS) $(INC) -shared -o $@ ref_wrap.cpp $(SRC)
clean:
	rm -rf _ref
