// ref_loaders.cpp — C entry points around the REFERENCE'S OWN volume loaders, compiled by oracle/ref.mk from where they lie under
// /root/reference: Source/VolumeTextureToolkit/Private/VolumeAsset/Loaders/MHDLoader.cpp (UMHDLoader::ParseVolumeInfoFromHeader,
// CreateVolumeFromFile ...) and VolumeLoader.cpp (IVolumeLoader::ReadFileAsString, LoadRawDataFileFromInfo, LoadAndConvertData, ConvertData),
// both unmodified. TEST INFRASTRUCTURE ONLY, part of oracle/_ref/libtbrm_ref.so.
//
// Ours in this file: the leaves those loaders call into the engine / the toolkit's UE-bound translation unit (TextureUtilities.cpp cannot be
// compiled: texture assets, packages, FCompression) — file reads with stdio, zlib's uncompress, the per-format dispatch onto the reference's
// own header templates (TextureUtilities.cpp:304-350), and texture "creation" that merely records format, size and bulk data.
#include <zlib.h>

#include <cstdio>

#include "VolumeAsset/Loaders/MHDLoader.h"
#include "TextureUtilities.h"

#include "../include/tbrm.h"

// ---- engine / toolkit leaves -----------------------------------------------------------------------------------------------------------
bool FFileHelper::LoadFileToString(FString& out, const TCHAR* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    out.s.clear();
    char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) out.s.append(buf, n);
    std::fclose(f);
    return true;
}
void FPaths::Split(const FString& full, FString& path, FString& name, FString& ext) {
    const size_t slash = full.s.find_last_of("/\\");
    path.s = slash == std::string::npos ? std::string() : full.s.substr(0, slash);
    std::string file = slash == std::string::npos ? full.s : full.s.substr(slash + 1);
    const size_t dot = file.find_last_of('.');
    name.s = dot == std::string::npos ? file : file.substr(0, dot);
    ext.s = dot == std::string::npos ? std::string() : file.substr(dot + 1);
}
UVolumeAsset* UVolumeAsset::CreateTransient(FString) { return new UVolumeAsset(); }
UVolumeAsset* UVolumeAsset::CreatePersistent(FString, const FString) { return new UVolumeAsset(); }

uint8* UVolumeTextureToolkit::LoadRawFileIntoArray(const FString FileName, const int64 ByteSize) {  // TextureUtilities.cpp:262-283
    FILE* f = std::fopen(*FileName, "rb");
    if (!f) return nullptr;
    uint8* data = new uint8[ByteSize];
    const size_t got = std::fread(data, 1, (size_t) ByteSize, f);
    std::fclose(f);
    if ((int64) got != ByteSize) {
        delete[] data;
        return nullptr;
    }
    return data;
}
uint8* UVolumeTextureToolkit::LoadZLibCompressedFileIntoArray(const FString FileName, const int64 UncompressedByteSize,
                                                              const int64 CompressedByteSize) {  // TextureUtilities.cpp:285-302
    uint8* packed = LoadRawFileIntoArray(FileName, CompressedByteSize);
    if (!packed) return nullptr;
    uint8* data = new uint8[UncompressedByteSize];
    uLongf n = (uLongf) UncompressedByteSize;
    const int rc = uncompress(data, &n, packed, (uLong) CompressedByteSize);  // FCompression::UncompressMemory(NAME_Zlib, ...)
    delete[] packed;
    if (rc != Z_OK) {
        delete[] data;
        return nullptr;
    }
    return data;
}
uint8* UVolumeTextureToolkit::NormalizeArrayByFormat(const EVolumeVoxelFormat VoxelFormat, uint8* InArray, const int64 ByteSize, float& OutInMin,
                                                     float& OutInMax) {  // TextureUtilities.cpp:304-327, the switch only
    switch (VoxelFormat) {
        case EVolumeVoxelFormat::UnsignedChar: return ConvertArrayToNormalizedArray<uint8, uint8>(InArray, ByteSize, OutInMin, OutInMax);
        case EVolumeVoxelFormat::SignedChar: return ConvertArrayToNormalizedArray<int8, uint8>(InArray, ByteSize, OutInMin, OutInMax);
        case EVolumeVoxelFormat::UnsignedShort: return ConvertArrayToNormalizedArray<uint16, uint16>(InArray, ByteSize, OutInMin, OutInMax);
        case EVolumeVoxelFormat::SignedShort: return ConvertArrayToNormalizedArray<int16, uint16>(InArray, ByteSize, OutInMin, OutInMax);
        case EVolumeVoxelFormat::UnsignedInt: return ConvertArrayToNormalizedArray<uint32, uint16>(InArray, ByteSize, OutInMin, OutInMax);
        case EVolumeVoxelFormat::SignedInt: return ConvertArrayToNormalizedArray<int32, uint16>(InArray, ByteSize, OutInMin, OutInMax);
        case EVolumeVoxelFormat::Float: return ConvertArrayToNormalizedArray<float, uint16>(InArray, ByteSize, OutInMin, OutInMax);
        default: return nullptr;
    }
}
float* UVolumeTextureToolkit::ConvertArrayToFloat(const EVolumeVoxelFormat VoxelFormat, uint8* InArray, uint64 VoxelCount) {  // :329-350
    switch (VoxelFormat) {
        case EVolumeVoxelFormat::UnsignedChar: return ConvertArrayToFloatTemplated<uint8>(InArray, VoxelCount);
        case EVolumeVoxelFormat::SignedChar: return ConvertArrayToFloatTemplated<int8>(InArray, VoxelCount);
        case EVolumeVoxelFormat::UnsignedShort: return ConvertArrayToFloatTemplated<uint16>(InArray, VoxelCount);
        case EVolumeVoxelFormat::SignedShort: return ConvertArrayToFloatTemplated<int16>(InArray, VoxelCount);
        case EVolumeVoxelFormat::UnsignedInt: return ConvertArrayToFloatTemplated<uint32>(InArray, VoxelCount);
        case EVolumeVoxelFormat::SignedInt: return ConvertArrayToFloatTemplated<int32>(InArray, VoxelCount);
        default: return nullptr;
    }
}
static int pixel_bytes(EPixelFormat f) { return f == PF_G8 ? 1 : (f == PF_G16 ? 2 : 4); }
static void record_texture(UVolumeTexture*& OutTexture, EPixelFormat PixelFormat, FIntVector Dimensions, uint8* BulkData) {
    OutTexture = new UVolumeTexture();
    OutTexture->PixelFormat = (int) PixelFormat;
    OutTexture->SizeX = Dimensions.X, OutTexture->SizeY = Dimensions.Y, OutTexture->SizeZ = Dimensions.Z;
    const size_t bytes = (size_t) Dimensions.X * Dimensions.Y * Dimensions.Z * pixel_bytes(PixelFormat);
    if (BulkData) OutTexture->Bulk.assign(BulkData, BulkData + bytes);
}
bool UVolumeTextureToolkit::CreateVolumeTextureTransient(UVolumeTexture*& OutTexture, EPixelFormat PixelFormat, FIntVector Dimensions, uint8* BulkData,
                                                         bool) {
    record_texture(OutTexture, PixelFormat, Dimensions, BulkData);
    return true;
}
bool UVolumeTextureToolkit::CreateVolumeTextureAsset(UVolumeTexture*& OutTexture, FString, FString, EPixelFormat PixelFormat, FIntVector Dimensions,
                                                     uint8* BulkData, bool, bool) {
    record_texture(OutTexture, PixelFormat, Dimensions, BulkData);
    return true;
}
void UVolumeTextureToolkit::SetupVolumeTexture(UVolumeTexture*& OutVolumeTexture, EPixelFormat PixelFormat, FIntVector Dimensions, uint8* InSourceArray,
                                               bool) {
    record_texture(OutVolumeTexture, PixelFormat, Dimensions, InSourceArray);
}

// ---- C entry points ------------------------------------------------------------------------------------------------------------------------
static void to_c(const FVolumeInfo& i, tbrm_volume_info* o) {
    std::memset(o, 0, sizeof(*o));
    o->parse_ok = i.bParseWasSuccessful ? 1 : 0;
    o->dims[0] = i.Dimensions.X, o->dims[1] = i.Dimensions.Y, o->dims[2] = i.Dimensions.Z;
    o->spacing[0] = i.Spacing.X, o->spacing[1] = i.Spacing.Y, o->spacing[2] = i.Spacing.Z;
    o->world_dims[0] = i.WorldDimensions.X, o->world_dims[1] = i.WorldDimensions.Y, o->world_dims[2] = i.WorldDimensions.Z;
    o->original_format = (int) i.OriginalFormat, o->actual_format = (int) i.ActualFormat;
    o->bytes_per_voxel = i.bParseWasSuccessful ? (int) i.BytesPerVoxel : 0;
    o->is_signed = i.bParseWasSuccessful ? (i.bIsSigned ? 1 : 0) : 0;
    o->is_normalized = i.bIsNormalized ? 1 : 0;
    o->min_value = i.MinValue, o->max_value = i.MaxValue;
    o->is_compressed = i.bIsCompressed ? 1 : 0;
    o->compressed_bytes = i.CompressedByteSize;
    std::snprintf(o->data_file, sizeof(o->data_file), "%s", i.DataFileName.s.c_str());
}

// UMHDLoader::ParseVolumeInfoFromHeader (MHDLoader.cpp:18-181) on a header FILE (the reference reads the file itself)
extern "C" int tbref_mhd_parse_file(const char* mhd_path, tbrm_volume_info* out) {
    UMHDLoader loader;
    to_c(loader.ParseVolumeInfoFromHeader(FString(mhd_path)), out);
    return out->parse_ok ? 0 : 1;
}

// UMHDLoader::CreateVolumeFromFile (MHDLoader.cpp:183-227): header, data file, IVolumeLoader::ConvertData, texture. Returns 0 and the asset's
// FVolumeInfo, the texture's pixel format as a tbrm_format (-1: a format the path does not sample) and its bulk data; 1 if no asset came out.
extern "C" int tbref_mhd_create_volume(const char* mhd_path, int normalize, int convert_to_float, tbrm_volume_info* out_info, int* out_format,
                                       void* out_data, uint64_t out_capacity, uint64_t* out_bytes) {
    UMHDLoader loader;
    UVolumeAsset* asset = loader.CreateVolumeFromFile(FString(mhd_path), normalize != 0, convert_to_float != 0);
    if (!asset || !asset->DataTexture) return 1;
    to_c(asset->ImageInfo, out_info);
    const EPixelFormat pf = (EPixelFormat) asset->DataTexture->PixelFormat;
    *out_format = pf == PF_G8 ? TBRM_FMT_G8 : (pf == PF_G16 ? TBRM_FMT_G16 : (pf == PF_R32_FLOAT ? TBRM_FMT_R32F : -1));
    *out_bytes = asset->DataTexture->Bulk.size();
    if (out_data && *out_bytes <= out_capacity) std::memcpy(out_data, asset->DataTexture->Bulk.data(), (size_t) *out_bytes);
    return 0;
}
