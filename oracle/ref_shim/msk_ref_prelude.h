// Stand-in for "misaki/core/fwd.h" (+ properties.h / transform.h / logger macros) when the reference's math headers are
// compiled on their own: fwd.h drags in the whole object system, the file resolver, bounding boxes and transforms.
// Everything numeric still comes from the reference's own headers, included from where they lie.
// TEST INFRASTRUCTURE (oracle/Makefile.ref); see oracle/ref_shim/Eigen/Core.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <limits>
#include <mutex>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
namespace std { using ::sinf; using ::cosf; } // mathutils.h:68-69 calls std::sinf / std::cosf, which libstdc++ does not declare
#include <misaki/core/platform.h>
#include <misaki/core/mathutils.h>
#include <misaki/core/frame.h>
#include <misaki/core/spectrum.h>
#include <misaki/core/distribution.h>
#include <misaki/core/properties.h> // the stand-in under oracle/ref_shim/misaki/core
#include "msk_ref_geometry.h"

namespace misaki {
using Distribution1D = math::Distribution1D<float>; // fwd.h:33-37
using Color3         = Color<float, 3>;
using Color4         = Color<float, 4>;
using Spectrum       = SpectrumArray<float, 4>;
using Wavelength     = SpectrumArray<float, 4>;

class Shape; class Emitter; class Scene; class Medium; class BSDF; class Texture; class Sampler; // fwd.h:41-60
class Sensor; class Film; class ImageBlock; class Integrator; class ReconstructionFilter; class Mesh;
namespace fs = std::filesystem; // fwd.h:39
class FileResolver { // fresolver.h: one search directory, MSK_REF_DATA_ROOT (for data/srgb.coeff); scene.cpp defines the global instance
public:
    fs::path resolve(const fs::path &p) const {
        const char *root = getenv("MSK_REF_DATA_ROOT");
        return (root && p.is_relative() && fs::exists(fs::path(root) / p)) ? fs::path(root) / p : p;
    }
};
FileResolver *get_file_resolver();
struct Ray; struct RayDifferential; struct PositionSample; struct DirectionSample; struct SceneInteraction; struct BSDFSample;
template <typename C> auto Properties::color(const std::string &n) const { const auto &c = colors.at(n); return Color3(c[0], c[1], c[2]); }
enum MskRefLogLevel { Trace, Debug, Info, Warn, Error };
inline const char *msk_ref_first() { return "?"; }
template <typename F, typename... A> inline const char *msk_ref_first(F &&f, A &&...) { return (const char *) f; }
template <typename... A> [[noreturn]] inline void msk_ref_throw(A &&...a) { fprintf(stderr, "[ref] Throw: %s\n", msk_ref_first(a...)); throw 1; }
template <typename... A> inline void msk_ref_log(A &&...) {}
} // namespace misaki
#define MSK_NOT_IMPLEMENTED(name) (fprintf(stderr, "[ref] not implemented: %s\n", name), throw 1)
namespace fmt { template <typename... A> inline std::string format(A &&...) { return std::string(); } } // used by to_string() only
#define Throw(...) ::misaki::msk_ref_throw(__VA_ARGS__)
#define Log(...) ::misaki::msk_ref_log(__VA_ARGS__)
