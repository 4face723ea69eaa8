// ingest.cu — volume ingest, the step right before the hot path, sm_100a.
//   UVolumeTextureToolkit::ConvertArrayToNormalizedArray / NormalizeArrayByFormat   Source/VolumeTextureToolkit/Public/TextureUtilities.h:103-149,
//                                                                                   Private/TextureUtilities.cpp:304-327
//   UVolumeTextureToolkit::ConvertArrayToFloat                                      TextureUtilities.h:153-178, TextureUtilities.cpp:329-350
//   UMHDLoader::ParseVolumeInfoFromHeader                                           Private/VolumeAsset/Loaders/MHDLoader.cpp:18-181
//   IVolumeLoader::LoadRawDataFileFromInfo / ConvertData                            Private/VolumeAsset/Loaders/VolumeLoader.cpp:16-128
// The reference scans the array twice on one CPU thread. Here: a min/max reduction (16-byte loads, warp shuffles, one partial per CTA) and
// a conversion pass whose CTAs first fold the partials — two launches, no atomics, no host round trip. HBM-bound byte work:
// algorithmic bytes per voxel = 2 * B_in + B_out (the second read comes from L2 when the volume fits its 126 MB).
#include <zlib.h>

#include <cstdio>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "tbrm_internal.hpp"

namespace tbrm {

constexpr int kIngestBlocks = 148 * 8;  // 8 resident CTAs of 256 threads per SM
constexpr int kIngestThreads = 256;

template <typename T, int N>
struct alignas(sizeof(T) * N >= 16 ? 16 : sizeof(T) * N) Pack {
    T v[N];
};

template <typename In>
struct MinMax {
    In mn, mx;
};

// Grid-stride walk over the 16-byte packs of an array with kInFlight loads of a thread issued before the first is consumed: 8 CTAs x 256
// threads x 16 bytes per SM keep 4.8 MB in flight with one load per thread — short of what 6.5 TB/s x the HBM latency needs; four do.
constexpr int kInFlight = 4;
template <typename In, int VEC, typename F>
__device__ __forceinline__ void for_each_pack(const In* __restrict__ in, size_t nvec, F&& body) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kInFlight - 1) * stride < nvec; i += kInFlight * stride) {
        Pack<In, VEC> p[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) p[u] = *reinterpret_cast<const Pack<In, VEC>*>(in + (i + u * stride) * VEC);
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) body(i + u * stride, p[u]);
    }
    for (; i < nvec; i += stride) body(i, *reinterpret_cast<const Pack<In, VEC>*>(in + i * VEC));
}

// InMin = numeric_limits<InType>::max(), InMax = numeric_limits<InType>::min() (TextureUtilities.h:110-111): for float, min() is the
// smallest POSITIVE normal — an all-negative float volume reports FLT_MIN as its maximum. Kept: parity with the reference.
template <typename In>
__device__ __forceinline__ MinMax<In> minmax_init() {
    MinMax<In> m;
    m.mn = std::numeric_limits<In>::max();
    m.mx = std::numeric_limits<In>::min();
    return m;
}
template <typename In>
__device__ __forceinline__ void minmax_add(MinMax<In>& m, In v) {  // NaN compares false both ways: never selected, like the reference
    if (v < m.mn) m.mn = v;
    if (v > m.mx) m.mx = v;
}
template <typename In>
__device__ __forceinline__ MinMax<In> minmax_block_reduce(MinMax<In> m, MinMax<In>* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const In a = __shfl_xor_sync(0xffffffffu, m.mn, o), b = __shfl_xor_sync(0xffffffffu, m.mx, o);
        if (a < m.mn) m.mn = a;
        if (b > m.mx) m.mx = b;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_part[warp] = m;
    __syncthreads();
    if (warp == 0) {
        m = lane < (int) (blockDim.x >> 5) ? s_part[lane] : minmax_init<In>();
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const In a = __shfl_xor_sync(0xffffffffu, m.mn, o), b = __shfl_xor_sync(0xffffffffu, m.mx, o);
            if (a < m.mn) m.mn = a;
            if (b > m.mx) m.mx = b;
        }
        if (lane == 0) s_part[0] = m;
    }
    __syncthreads();
    m = s_part[0];
    __syncthreads();
    return m;
}

// pass 1: one (min, max) partial per CTA. VEC elements per 16-byte load (1 = unaligned input, scalar loads).
template <typename In, int VEC>
__global__ void __launch_bounds__(kIngestThreads) ingest_minmax_kernel(const In* __restrict__ in, size_t n, MinMax<In>* __restrict__ partials) {
    __shared__ MinMax<In> s_part[kIngestThreads / 32];
    MinMax<In> m = minmax_init<In>();
    const size_t nvec = n / VEC;
    for_each_pack<In, VEC>(in, nvec, [&](size_t, const Pack<In, VEC>& p) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) minmax_add(m, p.v[k]);
    });
    if (blockIdx.x == 0)
        for (size_t i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) minmax_add(m, in[i]);  // tail
    m = minmax_block_reduce(m, s_part);
    if (threadIdx.x == 0) partials[blockIdx.x] = m;
}

// OutArray[i] = OutMin + (Normalized * (OutMax - OutMin)), Normalized = ((float) In[i] - InMin) / ((float) InMax - InMin): the float ->
// OutType conversion truncates (TextureUtilities.h:139-143); a NaN (constant volume: 0 / 0) converts to 0 as on x86
template <typename In, typename Out>
__device__ __forceinline__ Out normalize_one(In v, float fmn, float range, float omaxf) {
    const float nrm = ((float) v - fmn) / range;
    const float val = 0.0f + (nrm * omaxf);
    return (val >= 0.0f && val < omaxf + 1.0f) ? (Out) val : (Out) 0;
}

// pass 2: fold the partials (every CTA, redundantly: <= 1184 pairs), then convert
template <typename In, typename Out, int VEC>
__global__ void __launch_bounds__(kIngestThreads) ingest_normalize_kernel(const In* __restrict__ in, size_t n, const MinMax<In>* __restrict__ partials,
                                                                          int npartials, Out* __restrict__ out, float* __restrict__ out_minmax) {
    __shared__ MinMax<In> s_part[kIngestThreads / 32];
    MinMax<In> m = minmax_init<In>();
    for (int i = threadIdx.x; i < npartials; i += blockDim.x) {
        const MinMax<In> p = partials[i];
        if (p.mn < m.mn) m.mn = p.mn;
        if (p.mx > m.mx) m.mx = p.mx;
    }
    m = minmax_block_reduce(m, s_part);
    const float fmn = (float) m.mn, fmx = (float) m.mx;
    const float range = fmx - fmn;
    const float omaxf = (float) std::numeric_limits<Out>::max();
    if (blockIdx.x == 0 && threadIdx.x == 0) out_minmax[0] = fmn, out_minmax[1] = fmx;  // OutOriginalMin / OutOriginalMax
    const size_t nvec = n / VEC;
    if constexpr (sizeof(In) == 1) {
        // 1-byte voxels: the map has 256 entries — each thread evaluates the reference's expression for one input value, the voxels
        // then go through the table (same function on the same inputs: identical results, no division per voxel)
        static_assert(kIngestThreads == 256, "one table entry per thread");
        __shared__ Out lut[256];
        constexpr int kLo = (int) std::numeric_limits<In>::min();
        lut[threadIdx.x] = normalize_one<In, Out>((In) ((int) threadIdx.x + kLo), fmn, range, omaxf);
        __syncthreads();
        for_each_pack<In, VEC>(in, nvec, [&](size_t i, const Pack<In, VEC>& p) {
            Pack<Out, VEC> q;
#pragma unroll
            for (int k = 0; k < VEC; ++k) q.v[k] = lut[(int) p.v[k] - kLo];
            *reinterpret_cast<Pack<Out, VEC>*>(out + i * VEC) = q;
        });
        if (blockIdx.x == 0)
            for (size_t i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) out[i] = lut[(int) in[i] - kLo];
    } else {
        for_each_pack<In, VEC>(in, nvec, [&](size_t i, const Pack<In, VEC>& p) {
            Pack<Out, VEC> q;
#pragma unroll
            for (int k = 0; k < VEC; ++k) q.v[k] = normalize_one<In, Out>(p.v[k], fmn, range, omaxf);
            *reinterpret_cast<Pack<Out, VEC>*>(out + i * VEC) = q;
        });
        if (blockIdx.x == 0)
            for (size_t i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) out[i] = normalize_one<In, Out>(in[i], fmn, range, omaxf);
    }
}

// ConvertArrayToFloatTemplated: NewData[i] = static_cast<float>(TypedData[i])
template <typename In, int VEC>
__global__ void __launch_bounds__(kIngestThreads) ingest_to_float_kernel(const In* __restrict__ in, size_t n, float* __restrict__ out) {
    const size_t nvec = n / VEC;
    for_each_pack<In, VEC>(in, nvec, [&](size_t i, const Pack<In, VEC>& p) {
        Pack<float, VEC> q;
#pragma unroll
        for (int k = 0; k < VEC; ++k) q.v[k] = static_cast<float>(p.v[k]);
        *reinterpret_cast<Pack<float, VEC>*>(out + i * VEC) = q;
    });
    if (blockIdx.x == 0)
        for (size_t i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) out[i] = static_cast<float>(in[i]);
}

static int ingest_grid(size_t n, int vec) {
    const size_t want = (n / vec + kIngestThreads - 1) / kIngestThreads;
    return (int) std::max<size_t>(1, std::min<size_t>(want, kIngestBlocks));
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename In, typename Out>
static cudaError_t normalize_typed(cudaStream_t stream, const void* d_in, size_t n, void* d_out, void* d_partials, float* d_minmax) {
    constexpr int VEC = 16 / sizeof(In);
    auto* parts = (MinMax<In>*) d_partials;
    if (aligned16(d_in) && aligned16(d_out)) {
        const int grid = ingest_grid(n, VEC);
        ingest_minmax_kernel<In, VEC><<<grid, kIngestThreads, 0, stream>>>((const In*) d_in, n, parts);
        ingest_normalize_kernel<In, Out, VEC><<<grid, kIngestThreads, 0, stream>>>((const In*) d_in, n, parts, grid, (Out*) d_out, d_minmax);
    } else {
        const int grid = ingest_grid(n, 1);
        ingest_minmax_kernel<In, 1><<<grid, kIngestThreads, 0, stream>>>((const In*) d_in, n, parts);
        ingest_normalize_kernel<In, Out, 1><<<grid, kIngestThreads, 0, stream>>>((const In*) d_in, n, parts, grid, (Out*) d_out, d_minmax);
    }
    count_launch(2);
    return cudaGetLastError();
}

template <typename In>
static cudaError_t to_float_typed(cudaStream_t stream, const void* d_in, size_t n, float* d_out) {
    constexpr int VEC = 16 / sizeof(In);
    if (aligned16(d_in) && aligned16(d_out))
        ingest_to_float_kernel<In, VEC><<<ingest_grid(n, VEC), kIngestThreads, 0, stream>>>((const In*) d_in, n, d_out);
    else
        ingest_to_float_kernel<In, 1><<<ingest_grid(n, 1), kIngestThreads, 0, stream>>>((const In*) d_in, n, d_out);
    count_launch();
    return cudaGetLastError();
}

int voxel_format_bytes(int fmt) {  // FVolumeInfo::VoxelFormatByteSize — VolumeInfo.cpp:57-76
    switch (fmt) {
        case TBRM_VOXEL_U8:
        case TBRM_VOXEL_I8: return 1;
        case TBRM_VOXEL_U16:
        case TBRM_VOXEL_I16: return 2;
        case TBRM_VOXEL_U32:
        case TBRM_VOXEL_I32:
        case TBRM_VOXEL_F32: return 4;
        default: return 0;
    }
}
size_t ingest_partials_bytes() { return (size_t) kIngestBlocks * 8; }  // MinMax of a 4-byte type

// NormalizeArrayByFormat — TextureUtilities.cpp:304-327: 1-byte inputs normalise to u8, everything else to u16
cudaError_t ingest_normalize(cudaStream_t stream, int fmt, const void* d_in, size_t n, void* d_out, void* d_partials, float* d_minmax) {
    switch (fmt) {
        case TBRM_VOXEL_U8: return normalize_typed<uint8_t, uint8_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        case TBRM_VOXEL_I8: return normalize_typed<int8_t, uint8_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        case TBRM_VOXEL_U16: return normalize_typed<uint16_t, uint16_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        case TBRM_VOXEL_I16: return normalize_typed<int16_t, uint16_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        case TBRM_VOXEL_U32: return normalize_typed<uint32_t, uint16_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        case TBRM_VOXEL_I32: return normalize_typed<int32_t, uint16_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        case TBRM_VOXEL_F32: return normalize_typed<float, uint16_t>(stream, d_in, n, d_out, d_partials, d_minmax);
        default: return cudaErrorInvalidValue;
    }
}
// ConvertArrayToFloat — TextureUtilities.cpp:329-350 (a float input is rejected there, too)
cudaError_t ingest_to_float(cudaStream_t stream, int fmt, const void* d_in, size_t n, float* d_out) {
    switch (fmt) {
        case TBRM_VOXEL_U8: return to_float_typed<uint8_t>(stream, d_in, n, d_out);
        case TBRM_VOXEL_I8: return to_float_typed<int8_t>(stream, d_in, n, d_out);
        case TBRM_VOXEL_U16: return to_float_typed<uint16_t>(stream, d_in, n, d_out);
        case TBRM_VOXEL_I16: return to_float_typed<int16_t>(stream, d_in, n, d_out);
        case TBRM_VOXEL_U32: return to_float_typed<uint32_t>(stream, d_in, n, d_out);
        case TBRM_VOXEL_I32: return to_float_typed<int32_t>(stream, d_in, n, d_out);
        default: return cudaErrorInvalidValue;
    }
}

// ---- MetaImage header (host) ------------------------------------------------------------------------------------------------------------
// UMHDLoader::ParseVolumeInfoFromHeader reads whitespace-separated words: for each key it scans the words from the start of the file
// for the first one that EQUALS the key, skips the next word (the "=") and reads the values after it. Required keys, in this order:
// DimSize (3 ints), ElementSpacing or ElementSize (3 reals), ElementType (MET_*), ElementDataFile (one word); CompressedDataSize is
// optional and switches zlib loading on. A missing required key or an unknown element type fails the parse.
static bool words_after(const std::string& text, const char* key_a, const char* key_b, std::istringstream& in) {
    in = std::istringstream(text);
    std::string word;
    while (in >> word) {
        if (word == key_a || (key_b && word == key_b)) return static_cast<bool>(in >> word);  // the "=" sign
    }
    return false;
}

bool mhd_parse_header(const std::string& text, tbrm_volume_info& out) {
    std::memset(&out, 0, sizeof(out));
    out.min_value = -1000.0f, out.max_value = 3000.0f;  // FVolumeInfo defaults (VolumeInfo.h:103-109)
    std::istringstream in;
    if (!words_after(text, "DimSize", nullptr, in)) return false;
    in >> out.dims[0] >> out.dims[1] >> out.dims[2];
    if (in.fail() || out.dims[0] <= 0 || out.dims[1] <= 0 || out.dims[2] <= 0) return false;  // a DimSize that did not parse leaves zeros
    if (!words_after(text, "ElementSpacing", "ElementSize", in)) return false;
    in >> out.spacing[0] >> out.spacing[1] >> out.spacing[2];
    for (int k = 0; k < 3; ++k) out.world_dims[k] = out.spacing[k] * (double) out.dims[k];  // WorldDimensions = Spacing * Dimensions
    if (!words_after(text, "ElementType", nullptr, in)) return false;
    std::string type;
    in >> type;
    static const struct {
        const char* name;
        int fmt;
    } kTypes[] = {{"MET_UCHAR", TBRM_VOXEL_U8},   {"MET_CHAR", TBRM_VOXEL_I8},  {"MET_USHORT", TBRM_VOXEL_U16}, {"MET_SHORT", TBRM_VOXEL_I16},
                  {"MET_UINT", TBRM_VOXEL_U32},   {"MET_INT", TBRM_VOXEL_I32},  {"MET_FLOAT", TBRM_VOXEL_F32}};
    int fmt = -1;
    for (const auto& t : kTypes)
        if (type == t.name) fmt = t.fmt;
    if (fmt < 0) return false;
    out.original_format = fmt;
    out.actual_format = TBRM_VOXEL_U8;  // FVolumeInfo's default (VolumeInfo.h:72-74): ConvertData sets it once the data has been loaded
    out.bytes_per_voxel = voxel_format_bytes(fmt);
    out.is_signed = (fmt == TBRM_VOXEL_I8 || fmt == TBRM_VOXEL_I16 || fmt == TBRM_VOXEL_I32 || fmt == TBRM_VOXEL_F32) ? 1 : 0;  // IsVoxelFormatSigned
    if (words_after(text, "CompressedDataSize", nullptr, in)) {
        out.is_compressed = 1;
        in >> out.compressed_bytes;
        if (in.fail() || out.compressed_bytes <= 0) return false;
    }
    if (!words_after(text, "ElementDataFile", nullptr, in)) return false;
    std::string file;
    in >> file;
    std::snprintf(out.data_file, sizeof(out.data_file), "%s", file.c_str());
    out.parse_ok = 1;
    return true;
}

// LoadRawFileIntoArray / LoadZLibCompressedFileIntoArray (TextureUtilities.cpp:262-302): exactly `bytes` bytes of voxels
bool load_voxel_file(const std::string& path, const tbrm_volume_info& info, std::vector<uint8_t>& voxels, std::string& err) try {
    // the header values come from a file: reject what cannot be a volume before sizing anything by them
    constexpr long long kMaxSide = 1 << 16;
    constexpr unsigned long long kMaxBytes = 1ull << 40;
    if (info.dims[0] <= 0 || info.dims[1] <= 0 || info.dims[2] <= 0 || info.dims[0] > kMaxSide || info.dims[1] > kMaxSide || info.dims[2] > kMaxSide ||
        info.bytes_per_voxel <= 0 || info.bytes_per_voxel > 8) {
        err = "volume header: DimSize / ElementType out of range";
        return false;
    }
    const unsigned long long want = (unsigned long long) info.dims[0] * info.dims[1] * info.dims[2] * (unsigned long long) info.bytes_per_voxel;
    if (want > kMaxBytes || (info.is_compressed && (info.compressed_bytes <= 0 || (unsigned long long) info.compressed_bytes > kMaxBytes))) {
        err = "volume header: DimSize x ElementType (or CompressedDataSize) is not a plausible size";
        return false;
    }
    const size_t bytes = (size_t) want;
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        err = "cannot open " + path;
        return false;
    }
    if (info.is_compressed) {  // the file must at least hold what the header promises before it is sized by it
        f.seekg(0, std::ios::end);
        if ((long long) f.tellg() < (long long) info.compressed_bytes) {
            err = path + " holds fewer bytes than CompressedDataSize";
            return false;
        }
        f.seekg(0, std::ios::beg);
    }
    voxels.resize(bytes);
    if (!info.is_compressed) {
        f.read((char*) voxels.data(), (std::streamsize) bytes);
        if ((size_t) f.gcount() != bytes) {
            err = path + " holds fewer bytes than DimSize x ElementType";
            return false;
        }
        return true;
    }
    std::vector<uint8_t> packed((size_t) info.compressed_bytes);
    f.read((char*) packed.data(), (std::streamsize) packed.size());
    if ((size_t) f.gcount() != packed.size()) {
        err = path + " holds fewer bytes than CompressedDataSize";
        return false;
    }
    uLongf got = (uLongf) bytes;
    if (uncompress(voxels.data(), &got, packed.data(), (uLong) packed.size()) != Z_OK || got != bytes) {
        err = "zlib: " + path + " does not inflate to DimSize x ElementType bytes";
        return false;
    }
    return true;
} catch (const std::exception& e) {  // std::bad_alloc / std::length_error must not cross the C ABI
    err = std::string("loading ") + path + ": " + e.what();
    voxels.clear();
    return false;
}

}  // namespace tbrm
