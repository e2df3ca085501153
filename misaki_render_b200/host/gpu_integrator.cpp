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
#include <mutex>

namespace misaki {

class SamplingIntegrator : public Integrator {
public:
    MSK_DECLARE_CLASS()
protected:
    explicit SamplingIntegrator(const Properties &props) : Integrator(props) { // integrator.cpp:17-29
        m_block_size = (uint32_t) props.int_("block_size", 32);
        // <boolean> is dropped by the loader (as in the reference), so this is false unless set programmatically
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
struct DeviceContext { // one MskCtx per process and device, created on first use
    std::mutex mutex;
    MskCtx *ctx = nullptr;
    int device = -1;
    ~DeviceContext() { if (ctx) msk_gpu_shutdown(ctx); }
};
DeviceContext g_dev;
} // namespace

class GpuPathIntegrator final : public MonteCarloIntegrator {
public:
    explicit GpuPathIntegrator(const Properties &props) : MonteCarloIntegrator(props) {
        // The reference's PathTracer shadows m_max_depth / m_rr_depth with private members (-1 / 5), so the
        // XML values are silently ignored there (path.cpp:135-136).  They are honoured here: the
        // BASELINE configurations specify depths 5 and 16.
        m_device = (int) props.int_("device", -1);
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
    }

    bool render(Scene *scene, Sensor *sensor) override {
        Film *film = sensor->film();
        film->prepare({ "X", "Y", "Z", "A", "W" }); // integrator.cpp:36-41
        GpuSceneBuilder builder(scene);
        MskRenderDesc rd;
        render_desc(sensor, rd);
        int device = m_device >= 0 ? m_device : (getenv("MSK_DEVICE") ? atoi(getenv("MSK_DEVICE")) : 0);
        std::lock_guard<std::mutex> lock(g_dev.mutex);
        if (g_dev.ctx && g_dev.device != device) { msk_gpu_shutdown(g_dev.ctx); g_dev.ctx = nullptr; }
        if (!g_dev.ctx) {
            if (msk_gpu_init(device, &g_dev.ctx) != MSK_OK) Throw("%s", msk_gpu_last_error());
            g_dev.device = device;
        }
        MskScene *gpu_scene = nullptr;
        if (msk_gpu_scene_create(g_dev.ctx, &builder.desc(), &gpu_scene) != MSK_OK) Throw("%s", msk_gpu_last_error());
        MskAccelInfo info{};
        msk_gpu_accel_info(gpu_scene, &info);
        Log(Info, "GPU scene: %llu triangles, %llu wide nodes, BVH built in %.2f ms", (unsigned long long) info.ntris,
            (unsigned long long) info.nnodes, info.ms_build);
        Log(Info, "Start rendering...");
        ref<ImageBlock> block = new ImageBlock(film->width(), film->height(), 5);
        int rc = msk_gpu_render(gpu_scene, &rd, block->data().data(), &m_stats);
        msk_gpu_scene_destroy(gpu_scene);
        if (rc != MSK_OK) Throw("%s", msk_gpu_last_error());
        film->put(block.get()); // integrator.cpp:69 / hdrfilm.cpp:43-46
        double rays = (double) m_stats.rays_closest + (double) m_stats.rays_shadow;
        Log(Info, "Rendering finished. (took %.2f ms on the device: %.1f Mpaths/s, %.1f Mrays/s, %llu kernel launches)", m_stats.ms_render,
            m_stats.paths / (m_stats.ms_render * 1e3), rays / (m_stats.ms_render * 1e3), (unsigned long long) m_stats.kernel_launches);
        return true;
    }
    const MskStats &stats() const { return m_stats; }
    MSK_DECLARE_CLASS()
private:
    int m_device;
    int64_t m_sample_begin, m_sample_end;
    MskStats m_stats{};
};
MSK_IMPLEMENT_PLUGIN(GpuPathIntegrator, MonteCarloIntegrator, "path")

// used by the C API of the host library (capi.cpp)
bool gpu_path_render_desc(const Integrator *integrator, const Sensor *sensor, MskRenderDesc *rd) {
    auto *p = dynamic_cast<const GpuPathIntegrator *>(integrator);
    if (!p) return false;
    p->render_desc(sensor, *rd);
    return true;
}
bool gpu_path_stats(const Integrator *integrator, MskStats *stats) {
    auto *p = dynamic_cast<const GpuPathIntegrator *>(integrator);
    if (!p) return false;
    *stats = p->stats();
    return true;
}

} // namespace misaki
