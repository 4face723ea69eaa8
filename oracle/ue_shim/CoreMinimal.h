// CoreMinimal.h — SHIM (test infrastructure), not Unreal Engine code.
//
// Purpose: lets oracle/ref.mk compile a few UNMODIFIED source files of the reference plugin, from where they lie under
// /root/reference, into oracle/_ref/libtbrm_ref.so so that the oracle's host parameter math and volume normalisation can be
// checked against the reference's own code (tests/test_ref_pin_cpu.py). The reference needs Unreal Engine 5.4 to build; this
// header stands in for the handful of engine types those files touch. Everything here is OUR restatement of engine semantics
// (SURVEY.md Appendix B policies) — it pins the reference's logic (axis selection, offsets, step sizes, loop bounds,
// normalisation arithmetic), not the engine's.
//
// Engine semantics restated (UE 5.4, double-precision "large world coordinates"):
//   TVector<double>:  operator/=(s) multiplies by 1/s; Normalize(1e-8) scales by 1/sqrt(|v|^2) if |v|^2 > 1e-8; Size();
//   TVector2<double>: operator/=(s) multiplies by 1/s;
//   TQuat::UnrotateVector: v + w*t + q' x t with q' = -q.xyz, t = 2 (q' x v);
//   TTransform: InverseTransformVector = Unrotate(v) * SafeScaleReciprocal(scale), InverseTransformPosition = Unrotate(p - T) *
//               SafeScaleReciprocal(scale), InverseTransformVectorNoScale = Unrotate(v);
//   FLinearColor::ToFColor(bSRGB): clamp to [0,1], optional linear->sRGB, round to nearest byte (Q1/Q2).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

using std::abs;  // the reference calls unqualified abs() on doubles (MSVC resolves it to the double overload)

typedef uint8_t uint8;
typedef int8_t int8;
typedef uint16_t uint16;
typedef int16_t int16;
typedef uint32_t uint32;
typedef int32_t int32;
typedef uint64_t uint64;
typedef int64_t int64;
typedef char TCHAR;

#define UENUM(...)
#define USTRUCT(...)
#define UCLASS(...)
#define UPROPERTY(...)
#define UFUNCTION(...)
#define GENERATED_BODY()
#define TEXT(x) x
#define check(x) ((void) 0)
#define ensure(x) (x)
#define RAYMARCHER_API
#define VOLUMETEXTURETOOLKIT_API
#define DECLARE_LOG_CATEGORY_EXTERN(...)
#define DECLARE_MULTICAST_DELEGATE(...)
#define DECLARE_MULTICAST_DELEGATE_OneParam(...)
#define WITH_EDITOR 0
#define UE_SMALL_NUMBER (1.e-8)
#define UINTERFACE(...)
#define UE_LOG(...) ((void) 0)
#define DEFINE_LOG_CATEGORY(x)
#define TCHAR_TO_UTF8(x) (x)
#define MoveTemp(x) std::move(x)

enum class ESearchCase { CaseSensitive, IgnoreCase };
enum class ESearchDir { FromStart, FromEnd };
struct FString {
    std::string s;
    FString() {}
    FString(const char* c) : s(c) {}
    FString(const std::string& c) : s(c) {}
    static FString SanitizeFloat(double v) { return FString(std::to_string(v)); }
    const TCHAR* operator*() const { return s.c_str(); }
    int32 Find(const FString& sub, ESearchCase = ESearchCase::IgnoreCase, ESearchDir dir = ESearchDir::FromStart) const {
        const size_t p = dir == ESearchDir::FromEnd ? s.rfind(sub.s) : s.find(sub.s);
        return p == std::string::npos ? -1 : (int32) p;
    }
    void RightChopInline(int32 n) { s = n <= 0 ? s : (n >= (int32) s.size() ? std::string() : s.substr((size_t) n)); }
    void ReplaceCharInline(char from, char to) {
        for (char& c : s)
            if (c == from) c = to;
    }
    bool operator==(const FString& o) const { return s == o.s; }
};
inline FString operator+(const FString& a, const FString& b) { return FString(a.s + b.s); }
inline FString operator+(const char* a, const FString& b) { return FString(std::string(a) + b.s); }
inline FString operator+(const FString& a, const char* b) { return FString(a.s + std::string(b)); }

struct FIntPoint {
    int32 X = 0, Y = 0;
    FIntPoint() {}
    FIntPoint(int32 x, int32 y) : X(x), Y(y) {}
};
struct FIntVector {
    int32 X = 0, Y = 0, Z = 0;
    FIntVector() {}
    FIntVector(int32 x, int32 y, int32 z) : X(x), Y(y), Z(z) {}
    FString ToString() const { return FString("X=" + std::to_string(X) + " Y=" + std::to_string(Y) + " Z=" + std::to_string(Z)); }
};

struct FVector {
    double X, Y, Z;
    FVector() : X(0), Y(0), Z(0) {}
    FVector(double x, double y, double z) : X(x), Y(y), Z(z) {}
    explicit FVector(const FIntVector& v) : X(v.X), Y(v.Y), Z(v.Z) {}
    static double DotProduct(const FVector& a, const FVector& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
    static FVector CrossProduct(const FVector& a, const FVector& b) {
        return FVector(a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X);
    }
    FVector operator-() const { return FVector(-X, -Y, -Z); }
    FVector operator+(const FVector& o) const { return FVector(X + o.X, Y + o.Y, Z + o.Z); }
    FVector operator-(const FVector& o) const { return FVector(X - o.X, Y - o.Y, Z - o.Z); }
    FVector operator+(double b) const { return FVector(X + b, Y + b, Z + b); }
    FVector operator*(double s) const { return FVector(X * s, Y * s, Z * s); }
    FVector operator*(const FVector& o) const { return FVector(X * o.X, Y * o.Y, Z * o.Z); }
    FVector& operator*=(double s) {
        X *= s, Y *= s, Z *= s;
        return *this;
    }
    FVector& operator*=(const FVector& o) {
        X *= o.X, Y *= o.Y, Z *= o.Z;
        return *this;
    }
    FVector& operator/=(double s) {
        const double r = 1.0 / s;
        X *= r, Y *= r, Z *= r;
        return *this;
    }
    bool operator==(const FVector& o) const { return X == o.X && Y == o.Y && Z == o.Z; }
    double Size() const { return std::sqrt(X * X + Y * Y + Z * Z); }
    FString ToString() const { return FString("X=" + std::to_string(X) + " Y=" + std::to_string(Y) + " Z=" + std::to_string(Z)); }
    bool Normalize(double tol = UE_SMALL_NUMBER) {
        const double sq = X * X + Y * Y + Z * Z;
        if (sq > tol) {
            const double s = 1.0 / std::sqrt(sq);
            X *= s, Y *= s, Z *= s;
            return true;
        }
        return false;
    }
};
inline FVector operator*(double s, const FVector& v) { return v * s; }

struct FVector2D {
    double X, Y;
    FVector2D() : X(0), Y(0) {}
    FVector2D(double x, double y) : X(x), Y(y) {}
    FVector2D& operator/=(double s) {
        const double r = 1.0 / s;
        X *= r, Y *= r;
        return *this;
    }
};
struct FVector2f {
    float X, Y;
    FVector2f(float x, float y) : X(x), Y(y) {}
};

struct FQuat {
    double X = 0, Y = 0, Z = 0, W = 1;
    FVector UnrotateVector(const FVector& v) const {
        const FVector q(-X, -Y, -Z);
        const FVector t = 2.0 * FVector::CrossProduct(q, v);
        return v + (W * t) + FVector::CrossProduct(q, t);
    }
};

struct FTransform {
    FQuat Rotation;
    FVector Translation;
    FVector Scale3D = FVector(1, 1, 1);
    static FVector GetSafeScaleReciprocal(const FVector& s, double tol = UE_SMALL_NUMBER) {
        return FVector(std::fabs(s.X) <= tol ? 0.0 : 1.0 / s.X, std::fabs(s.Y) <= tol ? 0.0 : 1.0 / s.Y,
                       std::fabs(s.Z) <= tol ? 0.0 : 1.0 / s.Z);
    }
    FVector InverseTransformVector(const FVector& v) const { return Rotation.UnrotateVector(v) * GetSafeScaleReciprocal(Scale3D); }
    FVector InverseTransformVectorNoScale(const FVector& v) const { return Rotation.UnrotateVector(v); }
    FVector InverseTransformPosition(const FVector& p) const {
        return Rotation.UnrotateVector(p - Translation) * GetSafeScaleReciprocal(Scale3D);
    }
    FVector GetScale3D() const { return Scale3D; }
    bool Equals(const FTransform& o) const {
        return Translation == o.Translation && Scale3D == o.Scale3D && Rotation.X == o.Rotation.X && Rotation.Y == o.Rotation.Y &&
               Rotation.Z == o.Rotation.Z && Rotation.W == o.Rotation.W;
    }
};

struct FMatrix {
    double M[4][4];
    void SetIdentity() {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) M[i][j] = i == j ? 1.0 : 0.0;
    }
    // UE: axis i becomes ROW i of the matrix
    void SetAxes(const FVector* a0 = nullptr, const FVector* a1 = nullptr, const FVector* a2 = nullptr, const FVector* o = nullptr) {
        if (a0) M[0][0] = a0->X, M[0][1] = a0->Y, M[0][2] = a0->Z;
        if (a1) M[1][0] = a1->X, M[1][1] = a1->Y, M[1][2] = a1->Z;
        if (a2) M[2][0] = a2->X, M[2][1] = a2->Y, M[2][2] = a2->Z;
        if (o) M[3][0] = o->X, M[3][1] = o->Y, M[3][2] = o->Z;
    }
};

struct FColor {
    uint8 B, G, R, A;
    uint32 ToPackedARGB() const { return ((uint32) A << 24) | ((uint32) R << 16) | ((uint32) G << 8) | (uint32) B; }
};
struct FLinearColor {
    float R, G, B, A;
    FLinearColor() : R(0), G(0), B(0), A(0) {}
    FLinearColor(float r, float g, float b, float a = 1.0f) : R(r), G(g), B(b), A(a) {}
    static uint8 Quant(double c, bool srgb) {
        c = c < 0 ? 0 : (c > 1 ? 1 : c);  // NaN -> falls through as NaN; callers never pass NaN
        if (srgb) c = c <= 0.0031308 ? c * 12.92 : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055;
        return (uint8) std::floor(c * 255.0 + 0.5);
    }
    FColor ToFColor(bool srgb) const {
        FColor c;
        c.R = Quant(R, srgb), c.G = Quant(G, srgb), c.B = Quant(B, srgb), c.A = Quant(A, false);
        return c;
    }
};

enum EPixelFormat { PF_Unknown = 0, PF_G8, PF_G16, PF_R32_FLOAT, PF_R32_SINT, PF_R32_UINT, PF_FloatRGBA, PF_R16_SINT, PF_R16_UINT };
enum ETextureSourceFormat { TSF_Invalid = 0, TSF_G8, TSF_G16, TSF_RGBA16F };
enum TextureAddress { TA_Wrap = 0, TA_Clamp, TA_Mirror };

class UObject {
public:
    virtual ~UObject() {}
};
class UInterface : public UObject {};
class UDataAsset : public UObject {};
class UTexture : public UObject {};
class UTexture2D;
// what the loaders' texture-creation calls leave behind (ref_wrap.cpp fills it in): format, size and a copy of the bulk data
class UVolumeTexture : public UTexture {
public:
    int PixelFormat = 0;
    int32 SizeX = 0, SizeY = 0, SizeZ = 0;
    std::vector<uint8> Bulk;
};
struct FName {
    FString Name;
    FName() {}
    FName(const FString& n) : Name(n) {}
};
enum EObjectFlags { RF_NoFlags = 0, RF_Public = 1, RF_Standalone = 2 };
inline EObjectFlags operator|(EObjectFlags a, EObjectFlags b) { return (EObjectFlags) ((int) a | (int) b); }
template <typename T>
T* NewObject(UObject* = nullptr, FName = FName(), EObjectFlags = RF_NoFlags) { return new T(); }  // never freed: test process

template <typename T>
class TUniquePtr;
template <typename T>
class TUniquePtr<T[]> {
    std::unique_ptr<T[]> P;

public:
    TUniquePtr() {}
    explicit TUniquePtr(T* p) : P(p) {}
    TUniquePtr(TUniquePtr&&) = default;
    TUniquePtr& operator=(TUniquePtr&&) = default;
    T* Get() const { return P.get(); }
};
template <typename T>
struct TArray : std::vector<T> {
    void Empty() { this->clear(); }
};

// file-system stand-ins used by VolumeLoader.cpp (plain stdio; project-relative lookups resolve to nothing)
struct FFileHelper {
    static bool LoadFileToString(FString& out, const TCHAR* path);
};
struct FPaths {
    static void Split(const FString& full, FString& path, FString& name, FString& ext);
    static FString MakeValidFileName(const FString& n) { return n; }
    static FString ProjectContentDir() { return FString(""); }
    static bool DirectoryExists(const FString&) { return false; }
};
struct IFileManager {
    static IFileManager& Get() {
        static IFileManager m;
        return m;
    }
    FString ConvertToAbsolutePathForExternalAppForRead(const TCHAR* p) const { return FString(p); }
};
struct FFileManagerGeneric {
    static FFileManagerGeneric& Get() {
        static FFileManagerGeneric m;
        return m;
    }
    void FindFiles(TArray<FString>&, const TCHAR*, const TCHAR*) const {}
};
class UTextureRenderTargetVolume;
class UCurveLinearColor;
class URenderTargetVolumeMipped;

// task graph stand-ins for UVolumeTextureToolkit::ConvertArrayToFloatTemplated (TextureUtilities.h:153-178): the reference
// splits the array over the engine's worker threads; here the "workers" run one after the other (the result does not depend on it)
struct FTaskGraphInterface {
    static FTaskGraphInterface& Get() {
        static FTaskGraphInterface g;
        return g;
    }
    int32 GetNumWorkerThreads() const { return 4; }
};
template <typename F>
inline void ParallelFor(int32 n, F&& body) {
    for (int32 i = 0; i < n; ++i) body(i);
}
