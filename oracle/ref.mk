# Builds oracle/_ref/libtbrm_ref.so: the few host-side source files of the REFERENCE that can be compiled outside Unreal Engine,
# taken from where they lie under $(REF) (never copied into this repository), against the engine-type shim in oracle/ue_shim/
# plus our C wrapper (ref_wrap.cpp). Test infrastructure: used by tests/test_ref_pin_cpu.py and tests/golden/make_golden_ref.py
# to pin the oracle's host parameter math and the volume normalisation (row f3) to the reference's
# own code. The shaders (HLSL) and the RHI drivers cannot be built here; see DESIGN.md §2.
#     make -f ref.mk            (run from oracle/; build() of __graft_entry__.py does it when /root/reference exists)
REF ?= /root/reference
CXX := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
CXXFLAGS := -O2 -std=c++17 -fPIC -fopenmp -ffp-contract=off -fno-fast-math -mavx2 -mfma -w
INC := -I ue_shim -I $(REF)/Source/Raymarcher/Public -I $(REF)/Source/VolumeTextureToolkit/Public
SRC := $(REF)/Source/Raymarcher/Private/Rendering/LightingShaderUtils.cpp \
       $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/VolumeInfo.cpp \
       $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/MHDLoader.cpp \
       $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/VolumeLoader.cpp
SHADERS := $(REF)/Source/Raymarcher/Shaders/Private/AddDirLightShader.usf $(REF)/Source/Raymarcher/Shaders/Private/ChangeDirLightShader.usf \
           $(REF)/Source/Raymarcher/Shaders/Private/RaymarcherCommon.usf $(REF)/Source/Raymarcher/Shaders/Private/WindowedSampling.usf \
           $(REF)/Source/Raymarcher/Shaders/Private/RaymarchMaterialCommon.usf $(REF)/Source/Raymarcher/Shaders/Private/WindowedRaymarchMaterials.usf \
           $(REF)/Source/Raymarcher/Shaders/Private/GenerateOctreeShader.usf $(REF)/Source/FractalMarcher/Shaders/Private/SDFMarcher.usf \
           $(REF)/Source/FractalMarcher/Shaders/Private/CalculateMandelbulbSDF.usf
# two translation units: the engine-type shim (ue_shim, host C++) and the HLSL shim (hlsl_shim, shaders) define different worlds
_ref/libtbrm_ref.so: ref_wrap.cpp ref_loaders.cpp ref_shaders.cpp hlsl2cpp.py $(SRC) $(SHADERS) $(wildcard ue_shim/*.h) hlsl_shim/hlsl_shim.h tbrm_contract.h tbrm_oracle.h ../include/tbrm.h
	mkdir -p _ref/gen
	python3 hlsl2cpp.py $(REF) _ref/gen
	$(CXX) $(CXXFLAGS) $(INC) -c ref_wrap.cpp -o _ref/ref_wrap.o
	$(CXX) $(CXXFLAGS) $(INC) -c $(REF)/Source/Raymarcher/Private/Rendering/LightingShaderUtils.cpp -o _ref/LightingShaderUtils.o
	$(CXX) $(CXXFLAGS) $(INC) -c $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/VolumeInfo.cpp -o _ref/VolumeInfo.o
	# the loaders return a named rvalue-reference parameter by value (VolumeLoader.cpp:127): implicit move needs C++20 outside MSVC
	$(CXX) $(CXXFLAGS) -std=c++20 $(INC) -c $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/MHDLoader.cpp -o _ref/MHDLoader.o
	$(CXX) $(CXXFLAGS) -std=c++20 $(INC) -c $(REF)/Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/VolumeLoader.cpp -o _ref/VolumeLoader.o
	$(CXX) $(CXXFLAGS) -std=c++20 $(INC) -c ref_loaders.cpp -o _ref/ref_loaders.o
	$(CXX) $(CXXFLAGS) -I . -c ref_shaders.cpp -o _ref/ref_shaders.o
	$(CXX) $(CXXFLAGS) -shared -o $@ _ref/ref_wrap.o _ref/LightingShaderUtils.o _ref/VolumeInfo.o _ref/MHDLoader.o _ref/VolumeLoader.o _ref/ref_loaders.o _ref/ref_shaders.o -lz
	rm -rf _ref/gen _ref/*.o
clean:
	rm -rf _ref
