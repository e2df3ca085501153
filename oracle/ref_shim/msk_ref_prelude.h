// Stand-in for "misaki/core/fwd.h" (+ properties.h / transform.h / logger macros) when the reference's math headers are
// compiled on their own: fwd.h drags in the whole object system, the file resolver, bounding boxes and transforms.
// Everything numeric still comes from the reference's own headers, included from where they lie.
// TEST INFRASTRUCTURE (oracle/Makefile.ref); see oracle/ref_shim/Eigen/Core.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <tuple>
namespace std { using ::sinf; using ::cosf; } // mathutils.h:68-69 calls std::sinf / std::cosf, which libstdc++ does not declare
#include <misaki/core/platform.h>
#include <misaki/core/mathutils.h>
#include <misaki/core/frame.h>
#include <misaki/core/spectrum.h>
#include <misaki/core/distribution.h>

namespace misaki {
using Distribution1D = math::Distribution1D<float>; // fwd.h:33-37
using Color3         = Color<float, 3>;
using Color4         = Color<float, 4>;
using Spectrum       = SpectrumArray<float, 4>;
using Wavelength     = SpectrumArray<float, 4>;

// MicrofacetDistribution's first constructor reads a Properties object; the wrapper uses the (type, alpha_u, alpha_v)
// constructors, so an empty property set is enough for the header to compile
class Properties {
public:
    bool has_property(const std::string &) const { return false; }
    std::string string(const std::string &) const { return std::string(); }
    float float_(const std::string &) const { return 0.f; }
    bool bool_(const std::string &, bool def) const { return def; }
};
enum MskRefLogLevel { Trace, Debug, Info, Warn, Error };
template <typename... A> inline void msk_ref_throw(A &&...) { throw 1; }
template <typename... A> inline void msk_ref_log(A &&...) {}
} // namespace misaki
#define Throw(...) ::misaki::msk_ref_throw(__VA_ARGS__)
#define Log(...) ::misaki::msk_ref_log(__VA_ARGS__)
