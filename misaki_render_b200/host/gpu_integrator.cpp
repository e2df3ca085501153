// The "path" integrator plugin of the B200 backend: what a maintainer registers in place of
// reference src/librender/integrators/path.cpp (PathTracer, MSK_REGISTER_INSTANCE(PathTracer, "path") :140).
//
//   Integrator::render(Scene*, Sensor*)          reference include/misaki/render/integrator.h:9-17
//   SamplingIntegrator / MonteCarloIntegrator    reference src/librender/integrator.cpp:17-29,128-137 (parameters)
//
// render() flattens the object graph into the POD description of include/misaki_b200.h, calls the C ABI
// (msk_gpu_scene_create = Scene::accel_init, msk_gpu_render = the TBB tile loop + PathTracer::sample +
// ImageBlock::put), wraps the returned W x H x 5 XYZAW buffer in one border-less ImageBlock and hands it to
// Film::put -- the results contract of integrator.cpp:36-41,69.  A non-zero status becomes Throw(...), like
// every other error in the reference (logger.h:81-85).  There is no CPU fallback.
#include "render.h"

#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>

namespace misaki {

class SamplingIntegrator : public Integrator {
public:
    MSK_DECLARE_CLASS()
protected:
    explicit SamplingIntegrator(const Properties &props) : Integrator(props) { // integrator.cpp:17-29
        m_block_size = (uint32_t) props.int_("block_size", 32);
        m_hide_emitters = props.bool_("hide_emitters", false);
    }
    uint32_t m_block_size;
    bool m_hide_emitters;
};
MSK_IMPLEMENT_CLASS(SamplingIntegrator, Integrator)

class MonteCarloIntegrator : public SamplingIntegrator {
public:
    MSK_DECLARE_CLASS()
protected:
    explicit MonteCarloIntegrator(const Properties &props) : SamplingIntegrator(props) { // integrator.cpp:128-137
        m_rr_depth = (int) props.int_("rr_depth", 5);
        if (m_rr_depth <= 0) Throw("\"rr_depth\" must be set to a value greater than zero!");
        m_max_depth = (int) props.int_("max_depth", -1);
        if (m_max_depth < 0 && m_max_depth != -1) Throw("\"max_depth\" must be set to -1 (infinite) or a value >= 0");
    }
    int m_max_depth, m_rr_depth;
};
MSK_IMPLEMENT_CLASS(MonteCarloIntegrator, SamplingIntegrator)

namespace {
// One MskCtx per (device, ordinal) and process, created on first use.  The ordinal distinguishes repeated entries of a
// device list ("0,0": two contexts on one GPU, which is how the multi-device path is tested on a one-GPU box).
struct DevicePool {
    std::mutex mutex;
    std::map<std::pair<int, int>, MskCtx *> ctxs;
    MskCtx *get(int device, int ordinal) {
        auto key = std::make_pair(device, ordinal);
        auto it = ctxs.find(key);
        if (it != ctxs.end()) return it->second;
        MskCtx *ctx = nullptr;
        if (msk_gpu_init(device, &ctx) != MSK_OK) Throw("%s", msk_gpu_last_error());
        ctxs[key] = ctx;
        return ctx;
    }
    ~DevicePool() { for (auto &kv : ctxs) msk_gpu_shutdown(kv.second); }
};
DevicePool g_pool;

// The GPUs a render uses.  The reference uses every core of the machine from one Integrator::render call
// (integrator.cpp:54-75, main.cpp:60-61); here: the `devices` property of the integrator (a count; 0 = all visible GPUs),
// else the environment variable MSK_DEVICES (a count, "all", or a comma-separated list of device ids), else one GPU --
// `device` / MSK_DEVICE / 0, which is also the first GPU of a count.
std::vector<int> resolve_devices(int device_prop, int devices_prop) {
    const int first = device_prop >= 0 ? device_prop : (getenv("MSK_DEVICE") ? atoi(getenv("MSK_DEVICE")) : 0);
    int count = 1;
    if (devices_prop >= 0) count = devices_prop;
    else if (const char *env = getenv("MSK_DEVICES")) {
        std::string v(env);
        if (v.find(',') != std::string::npos) {
            std::vector<int> list;
            for (const std::string &tok : string::tokenize(v, ",")) list.push_back(atoi(tok.c_str()));
            if (list.empty()) Throw("MSK_DEVICES: empty device list");
            return list;
        }
        count = v == "all" ? 0 : atoi(v.c_str());
    }
    const int visible = msk_gpu_device_count();
    if (count <= 0) count = visible - first;
    if (count < 1 || first + count > visible) Throw("%d GPU(s) starting at device %d requested, %d visible", count, first, visible);
    std::vector<int> list;
    for (int i = 0; i < count; ++i) list.push_back(first + i);
    return list;
}
} // namespace

// Shared by the "path" and "aov" plugins: flatten the scene, run msk_gpu_render[_aov / _multi], hand the film-sized
// border-less block to Film::put.  `aov_types` empty and `aov == false`: the plain path tracer.  With several devices
// every GPU gets its own copy of the scene + BVH and a sample sub-range (msk_gpu_render_multi); the AOV integrator
// (one bounce) stays on the first.
static void gpu_render(Scene *scene, Sensor *sensor, const MskRenderDesc &rd, int device_prop, int devices_prop, bool aov,
                       const std::vector<int32_t> &aov_types, const std::vector<std::string> &aov_names, MskStats &stats) {
    Film *film = sensor->film();
    std::vector<std::string> channels = aov_names; // integrator.cpp:36-41: X,Y,Z,A,W first
    for (size_t i = 0; i < 5; ++i) channels.insert(channels.begin() + i, std::string(1, "XYZAW"[i]));
    film->prepare(channels);
    GpuSceneBuilder builder(scene);
    std::vector<int> devices = resolve_devices(device_prop, devices_prop);
    if (aov) devices.resize(1);
    std::lock_guard<std::mutex> lock(g_pool.mutex);
    const size_t nd = devices.size();
    std::vector<MskCtx *> ctxs(nd, nullptr);
    std::map<int, int> seen;
    for (size_t i = 0; i < nd; ++i) ctxs[i] = g_pool.get(devices[i], seen[devices[i]]++);
    // scene upload + BVH build, one host thread per GPU
    std::vector<MskScene *> gpu_scenes(nd, nullptr);
    std::vector<std::string> errors(nd);
    auto create = [&](size_t i) {
        if (msk_gpu_scene_create(ctxs[i], &builder.desc(), &gpu_scenes[i]) != MSK_OK) errors[i] = msk_gpu_last_error();
    };
    {
        std::vector<std::thread> threads;
        for (size_t i = 1; i < nd; ++i) threads.emplace_back(create, i);
        create(0);
        for (auto &t : threads) t.join();
    }
    auto destroy_scenes = [&]() { for (MskScene *s : gpu_scenes) msk_gpu_scene_destroy(s); };
    for (size_t i = 0; i < nd; ++i)
        if (!errors[i].empty()) { destroy_scenes(); Throw("%s", errors[i].c_str()); }
    MskAccelInfo info{};
    msk_gpu_accel_info(gpu_scenes[0], &info);
    Log(Info, "GPU scene: %llu triangles, %llu wide nodes, BVH built in %.2f ms (%zu GPU%s)", (unsigned long long) info.ntris,
        (unsigned long long) info.nnodes, info.ms_build, nd, nd == 1 ? "" : "s, samples partitioned");
    Log(Info, "Start rendering...");
    ref<ImageBlock> block = new ImageBlock(film->width(), film->height(), (uint32_t) channels.size());
    int rc;
    if (aov) {
        MskAovDesc ad{ aov_types.data(), (uint32_t) aov_types.size(), 0 };
        int nch = msk_gpu_aov_channels(&ad);
        if (nch < 0 || (size_t) nch + 5 != channels.size()) { destroy_scenes(); Throw("%s", nch < 0 ? msk_gpu_last_error() : "AOV channel names do not match the AOV types"); }
        rc = msk_gpu_render_aov(gpu_scenes[0], &rd, &ad, block->data().data(), &stats);
    } else if (nd > 1) {
        rc = msk_gpu_render_multi(gpu_scenes.data(), (uint32_t) nd, &rd, block->data().data(), &stats);
    } else {
        rc = msk_gpu_render(gpu_scenes[0], &rd, block->data().data(), &stats);
    }
    std::string err = rc != MSK_OK ? msk_gpu_last_error() : "";
    destroy_scenes();
    if (rc != MSK_OK) Throw("%s", err.c_str());
    film->put(block.get()); // integrator.cpp:69 / hdrfilm.cpp:43-46
    double rays = (double) stats.rays_closest + (double) stats.rays_shadow;
    Log(Info, "Rendering finished. (took %.2f ms on the device%s: %.1f Mpaths/s, %.1f Mrays/s, %llu kernel launches)", stats.ms_render,
        nd == 1 ? "" : "s", stats.paths / (stats.ms_render * 1e3), rays / (stats.ms_render * 1e3), (unsigned long long) stats.kernel_launches);
}

class GpuPathIntegrator : public MonteCarloIntegrator {
public:
    explicit GpuPathIntegrator(const Properties &props) : MonteCarloIntegrator(props) {
        // The reference's PathTracer shadows m_max_depth / m_rr_depth with private members (-1 / 5), so the
        // XML values are silently ignored there (path.cpp:135-136).  They are honoured here: the
        // BASELINE configurations specify depths 5 and 16.
        m_device = (int) props.int_("device", -1);
        m_devices = (int) props.int_("devices", -1); // number of GPUs (0 = all visible); default: MSK_DEVICES, else 1
        m_sample_begin = props.int_("sample_begin", 0);
        m_sample_end = props.int_("sample_end", -1);
    }

    void render_desc(const Sensor *sensor, MskRenderDesc &rd) const {
        rd = MskRenderDesc{};
        rd.spp = sensor->sampler()->sample_count();
        rd.sample_begin = (uint32_t) m_sample_begin;
        rd.sample_end = m_sample_end < 0 ? rd.spp : (uint32_t) m_sample_end;
        rd.max_depth = m_max_depth; rd.rr_depth = m_rr_depth; rd.hide_emitters = m_hide_emitters;
        rd.base_seed = sensor->sampler()->base_seed();
        rd.clear_film = 1;
        rd.integrator = m_integrator;
    }

    bool render(Scene *scene, Sensor *sensor) override {
        MskRenderDesc rd;
        render_desc(sensor, rd);
        gpu_render(scene, sensor, rd, m_device, m_devices, false, {}, {}, m_stats);
        return true;
    }
    const MskStats &stats() const { return m_stats; }
    int device() const { return m_device; }
    MSK_DECLARE_CLASS()
protected:
    uint32_t m_integrator = MSK_INTEGRATOR_PATH;
private:
    int m_device, m_devices;
    int64_t m_sample_begin, m_sample_end;
    MskStats m_stats{};
};
MSK_IMPLEMENT_PLUGIN(GpuPathIntegrator, MonteCarloIntegrator, "path")

// The "volpath" integrator plugin: reference src/librender/integrators/volpath.cpp (VolumetricPathTracer,
// MSK_REGISTER_INSTANCE(VolumetricPathTracer, "volpath") :182) -- same parameters and render() contract, the
// wavefront runs k_shade_vol (homogeneous media + isotropic phase function) instead of the path tracer's k_shade.
class GpuVolPathIntegrator final : public GpuPathIntegrator {
public:
    explicit GpuVolPathIntegrator(const Properties &props) : GpuPathIntegrator(props) { m_integrator = MSK_INTEGRATOR_VOLPATH; }
    MSK_DECLARE_CLASS()
};
MSK_IMPLEMENT_PLUGIN(GpuVolPathIntegrator, GpuPathIntegrator, "volpath")

// The "aov" integrator plugin: reference src/librender/integrators/aov.cpp (AOVIntegrator,
// MSK_REGISTER_INSTANCE(AOVIntegrator, "aov") :156).  The constructor follows aov.cpp:30-85: the "aovs" string is a
// list of <name>:<type> pairs (depth, position, uv, geo_normal, sh_normal), child integrators add
// <child>.R/.G/.B/.A.  sample() (aov.cpp:87-144) runs on the device behind msk_gpu_render_aov; one nested
// integrator, the GPU path tracer, is supported (nested integrators of other kinds do not exist in this build).
class GpuAovIntegrator final : public MonteCarloIntegrator {
public:
    explicit GpuAovIntegrator(const Properties &props) : MonteCarloIntegrator(props) {
        m_device = (int) props.int_("device", -1);
        for (const std::string &token : string::tokenize(props.string("aovs"))) {
            std::vector<std::string> item = string::tokenize(token, ":");
            if (item.size() != 2 || item[0].empty() || item[1].empty()) {
                Log(Warn, "Invalid AOV specification: require <name>:<type> pair");
                continue; // the reference goes on to read item[1] out of range here (aov.cpp:37-41)
            }
            if (item[1] == "depth") { m_types.push_back(MSK_AOV_DEPTH); m_names.push_back(item[0]); }
            else if (item[1] == "position") { m_types.push_back(MSK_AOV_POSITION); add3(item[0], "XYZ"); }
            else if (item[1] == "uv") { m_types.push_back(MSK_AOV_UV); m_names.push_back(item[0] + ".U"); m_names.push_back(item[0] + ".V"); }
            else if (item[1] == "geo_normal") { m_types.push_back(MSK_AOV_GEO_NORMAL); add3(item[0], "XYZ"); }
            else if (item[1] == "sh_normal") { m_types.push_back(MSK_AOV_SH_NORMAL); add3(item[0], "XYZ"); }
            else Throw("Invalid AOV type \"%s\"!", item[1].c_str());
        }
        for (auto &[name, obj] : props.objects()) {
            auto *path = dynamic_cast<GpuPathIntegrator *>(obj.get());
            if (!path) Throw("Child objects must be of type 'SamplingIntegrator'!");
            if (m_nested) Throw("The GPU AOV integrator supports one nested integrator");
            m_nested = path;
            m_types.push_back(MSK_AOV_INTEGRATOR_RGBA);
            for (const char *c : { ".R", ".G", ".B", ".A" }) m_names.push_back(name + c);
        }
        if (m_names.empty()) Log(Warn, "No AOVs were specified!");
    }
    bool render(Scene *scene, Sensor *sensor) override {
        MskRenderDesc rd{};
        if (m_nested) m_nested->render_desc(sensor, rd); // the nested tracer's depth / Russian-roulette parameters
        else {
            rd.spp = sensor->sampler()->sample_count(); rd.sample_end = rd.spp;
            rd.max_depth = m_max_depth; rd.rr_depth = m_rr_depth; rd.hide_emitters = m_hide_emitters;
            rd.base_seed = sensor->sampler()->base_seed(); rd.clear_film = 1;
        }
        gpu_render(scene, sensor, rd, m_device >= 0 ? m_device : (m_nested ? m_nested->device() : -1), 1, true, m_types, m_names, m_stats);
        return true;
    }
    const MskStats &stats() const { return m_stats; }
    const std::vector<std::string> &aov_names() const { return m_names; }
    MSK_DECLARE_CLASS()
private:
    void add3(const std::string &base, const char *suffix) {
        for (int i = 0; i < 3; ++i) m_names.push_back(base + "." + suffix[i]);
    }
    int m_device;
    std::vector<int32_t> m_types;
    std::vector<std::string> m_names;
    ref<GpuPathIntegrator> m_nested;
    MskStats m_stats{};
};
MSK_IMPLEMENT_PLUGIN(GpuAovIntegrator, MonteCarloIntegrator, "aov")

// used by the C API of the host library (capi.cpp)
bool gpu_path_render_desc(const Integrator *integrator, const Sensor *sensor, MskRenderDesc *rd) {
    auto *p = dynamic_cast<const GpuPathIntegrator *>(integrator);
    if (!p) return false;
    p->render_desc(sensor, *rd);
    return true;
}
bool gpu_path_stats(const Integrator *integrator, MskStats *stats) {
    if (auto *p = dynamic_cast<const GpuPathIntegrator *>(integrator)) { *stats = p->stats(); return true; }
    if (auto *a = dynamic_cast<const GpuAovIntegrator *>(integrator)) { *stats = a->stats(); return true; }
    return false;
}

} // namespace misaki
