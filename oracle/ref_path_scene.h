// The scene handle shared by ref_path_wrap.cpp (which owns PathTracer / AreaLight / ...) and ref_render_wrap.cpp.
// TEST INFRASTRUCTURE.
#pragma once
#include "ref_wrap_common.h"
#include <misaki/render/integrator.h>
#include <misaki/render/scene.h>
namespace misaki {
class RefScene final : public Scene {
public:
    explicit RefScene(const Properties &props) : Scene(props) {}
    std::string to_string() const override { return "RefScene"; }
};
} // namespace misaki
struct RefPathScene {
    std::vector<misaki::RefMesh *> meshes;
    misaki::RefScene *scene = nullptr;
    misaki::SamplingIntegrator *tracer = nullptr;
};
