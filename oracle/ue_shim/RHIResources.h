// SHIM (see CoreMinimal.h in this directory): just enough of the RHI for the reference's headers to parse.
#pragma once
#include "CoreMinimal.h"
class FRHITexture {};
class FRHITexture3D : public FRHITexture {
public:
    uint32 SizeX = 0, SizeY = 0, SizeZ = 0;
    uint32 GetSizeX() const { return SizeX; }
    uint32 GetSizeY() const { return SizeY; }
    uint32 GetSizeZ() const { return SizeZ; }
};
class FRHIUnorderedAccessView {};
class FRHICommandListImmediate;
struct FTexture2DRHIRef {};
struct FTexture3DRHIRef {};
struct FUnorderedAccessViewRHIRef {};
struct FSamplerStateRHIRef {
    uint32 BorderColor = 0;
};
enum ESamplerFilter { SF_Point, SF_Bilinear, SF_Trilinear };
enum ESamplerAddressMode { AM_Wrap, AM_Clamp, AM_Mirror, AM_Border };
struct FSamplerStateInitializerRHI {
    uint32 BorderColor;
    FSamplerStateInitializerRHI(ESamplerFilter, ESamplerAddressMode, ESamplerAddressMode, ESamplerAddressMode, float, int32, float, float,
                                uint32 InBorderColor)
        : BorderColor(InBorderColor) {}
};
inline FSamplerStateRHIRef RHICreateSamplerState(const FSamplerStateInitializerRHI& i) { return FSamplerStateRHIRef{i.BorderColor}; }
