// C entry points of the host library (libmisaki_host.so) -- a thin handle API over xml::load_file,
// Integrator::render and Film::develop so that tests and other languages can drive the C++ front-end.
#include "host_capi.h"
#include "render.h"

#include <cstring>

using namespace misaki;

namespace misaki {
bool gpu_path_render_desc(const Integrator *integrator, const Sensor *sensor, MskRenderDesc *rd);
bool gpu_path_stats(const Integrator *integrator, MskStats *stats);
}

struct MskhScene {
    ref<Object> root;
    Scene *scene = nullptr;
    std::unique_ptr<GpuSceneBuilder> builder;
};

static thread_local std::string g_host_error;

template <typename F> static int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return -1;
    }
}

static void finish_load(MskhScene *h) {
    h->scene = dynamic_cast<Scene *>(h->root.get());
    if (!h->scene) Throw("Root element of the input file must be a <scene> tag!"); // main.cpp:21-23
    if (!h->scene->integrator()) Throw("No integrator specified for scene");
}

extern "C" {

const char *mskh_last_error(void) { return g_host_error.c_str(); }
void mskh_set_log_level(int level) { set_log_level((LogLevel) level); }
void mskh_add_search_path(const char *dir) { get_file_resolver()->append(dir); }

int mskh_load_file_params(const char *filename, const char *const *names, const char *const *values, size_t nparams, MskhScene **out) {
    *out = nullptr;
    return guarded([&] {
        xml::ParameterList params;
        for (size_t i = 0; i < nparams; ++i) params.emplace_back(names[i], values[i]);
        std::string path = filename;
        size_t slash = path.find_last_of('/');
        ScopedSearchPath scene_dir(slash == std::string::npos ? "." : path.substr(0, slash)); // main.cpp:68, for this load only
        std::unique_ptr<MskhScene> h(new MskhScene);
        h->root = xml::load_file(get_file_resolver()->resolve(path), params);
        finish_load(h.get());
        *out = h.release();
    });
}

int mskh_load_file(const char *filename, MskhScene **out) { return mskh_load_file_params(filename, nullptr, nullptr, 0, out); }

int mskh_load_string(const char *xml_text, const char *base_dir, MskhScene **out) {
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<ScopedSearchPath> scene_dir;
        if (base_dir && *base_dir) scene_dir.reset(new ScopedSearchPath(base_dir));
        std::unique_ptr<MskhScene> h(new MskhScene);
        h->root = xml::load_string(xml_text);
        finish_load(h.get());
        *out = h.release();
    });
}

void mskh_free(MskhScene *h) { delete h; }

const MskSceneDesc *mskh_scene_desc(MskhScene *h) {
    if (guarded([&] { if (!h->builder) h->builder.reset(new GpuSceneBuilder(h->scene)); }) != 0) return nullptr;
    return &h->builder->desc();
}

int mskh_render_desc(MskhScene *h, MskRenderDesc *rd) {
    return guarded([&] {
        if (!gpu_path_render_desc(h->scene->integrator(), h->scene->sensor(), rd)) Throw("the scene's integrator is not the GPU path tracer");
    });
}

int mskh_render(MskhScene *h, const char *output_filename, MskStats *stats) {
    return guarded([&] {
        Scene *scene = h->scene;
        Sensor *sensor = scene->sensor();
        if (!sensor) Throw("The scene has no sensor");
        Film *film = sensor->film();
        if (output_filename && *output_filename) film->set_destination_file(output_filename);
        bool ok = scene->integrator()->render(scene, sensor); // main.cpp:38
        if (!ok) Throw("Rendering failed, result not saved.");
        if (output_filename && *output_filename) film->develop(); // main.cpp:40
        if (stats) gpu_path_stats(scene->integrator(), stats);
    });
}

int mskh_develop(const float *film_xyzaw, size_t npixels, float *rgba) {
    return guarded([&] { develop_xyzaw(film_xyzaw, npixels, rgba); });
}
int mskh_develop_channels(const float *film, size_t npixels, size_t nchannels, float *out) {
    return guarded([&] {
        if (nchannels < 5) Throw("develop: expected at least the X, Y, Z, A, W channels");
        develop_channels(film, npixels, nchannels, out);
    });
}
int mskh_write_exr(const char *filename, const float *rgba, uint32_t width, uint32_t height) {
    return guarded([&] { write_exr_rgba(filename, rgba, width, height); });
}
int mskh_write_pfm(const char *filename, const float *rgba, uint32_t width, uint32_t height) {
    return guarded([&] { write_pfm_rgb(filename, rgba, width, height); });
}
int mskh_srgb_model_fetch(const float rgb[3], float out[3]) {
    return guarded([&] { Color3 c = srgb_model_fetch(Color3{ rgb[0], rgb[1], rgb[2] }); out[0] = c.r; out[1] = c.g; out[2] = c.b; });
}

int mskh_registered_plugins(char *buffer, size_t size) {
    std::string all;
    for (auto &n : InstanceManager::get()->registered()) { all += n; all += ' '; }
    if (buffer && size) { strncpy(buffer, all.c_str(), size - 1); buffer[size - 1] = '\0'; }
    return (int) all.size();
}

} // extern "C"
