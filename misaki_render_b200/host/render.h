// Render-side plugin interfaces of the host front-end: the classes a misaki scene file instantiates
// (reference include/misaki/render/*.h).  The virtual per-sample methods of the reference (BSDF::sample,
// Emitter::eval, Texture::eval, Scene::ray_intersect ...) run on the GPU in this build, so the host classes
// carry the plugin PARAMETERS and know how to describe themselves to the device through GpuSceneBuilder;
// construction-time behaviour (defaults, validation, error messages, object ordering) follows the reference.
#pragma once
#include "core.h"
#include "../../include/misaki_b200.h"

namespace misaki {

class GpuSceneBuilder;
class Shape;
class Scene;
class Sensor;

// ---- Texture / spectra (reference include/misaki/render/texture.h, src/librender/spectra/*.cpp)
class Texture : public Object {
public:
    // appends this spectrum to the device description and returns its id; `unbounded`: conductor eta / k
    virtual int describe(GpuSceneBuilder &b) const = 0;
    virtual float mean() const { return 0.f; }
    static ref<Texture> D65(float scale); // texture.cpp:26-37
    MSK_DECLARE_CLASS()
protected:
    explicit Texture(const Properties &props) : m_id(props.id()) {}
    std::string m_id;
};

// ---- BSDF (reference include/misaki/render/bsdf.h)
class BSDF : public Object {
public:
    virtual int describe(GpuSceneBuilder &b) const = 0; // returns the bsdf id
    virtual bool has_transmission() const { return false; }
    MSK_DECLARE_CLASS()
protected:
    explicit BSDF(const Properties &props) : m_id(props.id()) {}
    std::string m_id;
};

// ---- Emitter (reference include/misaki/render/emitter.h)
class Emitter : public Object {
public:
    virtual bool is_environment() const { return false; }
    virtual bool is_surface() const { return false; }
    virtual MskEmitterType gpu_type() const = 0;
    const Texture *radiance() const { return m_radiance.get(); }
    void set_shape(Shape *s) { m_shape = s; }
    Shape *shape() const { return m_shape; }
    MSK_DECLARE_CLASS()
protected:
    explicit Emitter(const Properties &props);
    ref<Texture> m_radiance;
    Shape *m_shape = nullptr;
    Transform4f m_world_transform;
};

// ---- PhaseFunction / Medium (reference include/misaki/render/phase.h, medium.h; src/librender/phase.cpp, medium.cpp)
class PhaseFunction : public Object {
public:
    virtual MskPhaseType gpu_type() const = 0;
    MSK_DECLARE_CLASS()
protected:
    explicit PhaseFunction(const Properties &props) : m_id(props.id()) {}
    std::string m_id;
};

class Medium : public Object {
public:
    const PhaseFunction *phase_function() const { return m_phase_function.get(); }
    virtual void describe(GpuSceneBuilder &b, MskMedium &out) const = 0;
    MSK_DECLARE_CLASS()
protected:
    explicit Medium(const Properties &props); // medium.cpp:13-31
    ref<PhaseFunction> m_phase_function;
    std::string m_id;
};

// ---- Shape / Mesh (reference include/misaki/render/shape.h, mesh.h)
class Shape : public Object {
public:
    bool is_medium_transition() const { return m_interior_medium || m_exterior_medium; } // shape.h:35-38
    const Medium *interior_medium() const { return m_interior_medium.get(); }
    const Medium *exterior_medium() const { return m_exterior_medium.get(); }
    bool is_emitter() const { return (bool) m_emitter; }
    Emitter *emitter() const { return m_emitter.get(); }
    const BSDF *bsdf() const { return m_bsdf.get(); }
    void set_children();
    MSK_DECLARE_CLASS()
protected:
    explicit Shape(const Properties &props);
    ref<BSDF> m_bsdf;
    ref<Emitter> m_emitter;
    ref<Medium> m_interior_medium, m_exterior_medium;
    Transform4f m_world_transform;
    std::string m_id;
};

class Mesh : public Shape {
public:
    uint32_t vertex_count() const { return m_vertex_count; }
    uint32_t face_count() const { return m_face_count; }
    const float *vertices() const { return m_vertices.data(); } // 8 floats per vertex: p n uv, mesh.h / obj.cpp:139-142
    const uint32_t *faces() const { return m_faces.data(); }
    bool has_vertex_normals() const { return m_normal_offset != 0; }
    bool has_vertex_texcoords() const { return m_texcoord_offset != 0; }
    const std::string &name() const { return m_name; }
    MSK_DECLARE_CLASS()
protected:
    explicit Mesh(const Properties &props);
    std::string m_name;
    Transform4f m_to_world;
    std::vector<float> m_vertices;
    std::vector<uint32_t> m_faces;
    uint32_t m_vertex_count = 0, m_face_count = 0, m_normal_offset = 0, m_texcoord_offset = 0;
};

// ---- ReconstructionFilter (reference include/misaki/render/rfilter.h, src/librender/rfilter.cpp)
constexpr int MSK_FILTER_RESOLUTION = 32;
class ReconstructionFilter : public Object {
public:
    virtual float eval(float x) const = 0;
    float radius() const { return m_radius; }
    uint32_t border_size() const { return m_border_size; }
    const std::vector<float> &values() const { return m_values; }
    MSK_DECLARE_CLASS()
protected:
    explicit ReconstructionFilter(const Properties &) {}
    void init_discretization(); // rfilter.cpp:12-27
    std::vector<float> m_values;
    float m_radius = 0.f, m_scale_factor = 0.f;
    uint32_t m_border_size = 0;
};

// ---- ImageBlock (reference include/misaki/render/imageblock.h): here only the border-less whole-film block
// the GPU integrator hands to Film::put
class ImageBlock : public Object {
public:
    ImageBlock(uint32_t width, uint32_t height, uint32_t channels) : m_width(width), m_height(height), m_channels(channels),
        m_data((size_t) width * height * channels, 0.f) {}
    uint32_t width() const { return m_width; }
    uint32_t height() const { return m_height; }
    uint32_t channel_count() const { return m_channels; }
    std::vector<float> &data() { return m_data; }
    const std::vector<float> &data() const { return m_data; }
    MSK_DECLARE_CLASS()
private:
    uint32_t m_width, m_height, m_channels;
    std::vector<float> m_data;
};

// ---- Film (reference include/misaki/render/film.h, src/librender/film.cpp, films/hdrfilm.cpp)
class Film : public Object {
public:
    virtual void prepare(const std::vector<std::string> &channels) = 0;
    virtual void put(const ImageBlock *block) = 0;
    virtual void develop() = 0;
    virtual void set_destination_file(const std::string &filename) = 0;
    uint32_t width() const { return m_width; }
    uint32_t height() const { return m_height; }
    const ReconstructionFilter *filter() const { return m_filter.get(); }
    MSK_DECLARE_CLASS()
protected:
    explicit Film(const Properties &props);
    uint32_t m_width, m_height;
    ref<ReconstructionFilter> m_filter;
};

// ---- Sampler (reference include/misaki/render/sampler.h, src/librender/sampler.cpp)
class Sampler : public Object {
public:
    uint32_t sample_count() const { return m_sample_count; }
    uint64_t base_seed() const { return m_base_seed; }
    MSK_DECLARE_CLASS()
protected:
    explicit Sampler(const Properties &props);
    uint32_t m_sample_count;
    uint64_t m_base_seed;
};

// ---- Sensor (reference include/misaki/render/sensor.h, src/librender/sensor.cpp)
class Sensor : public Object {
public:
    Film *film() const { return m_film.get(); }
    Sampler *sampler() const { return m_sampler.get(); }
    const Medium *medium() const { return m_medium.get(); } // sensor.h:45-46
    virtual void describe(MskCamera &cam) const = 0;
    MSK_DECLARE_CLASS()
protected:
    explicit Sensor(const Properties &props);
    Transform4f m_world_transform;
    ref<Film> m_film;
    ref<Sampler> m_sampler;
    ref<Medium> m_medium;
    float m_aspect = 1.f;
};

// ---- Integrator (reference include/misaki/render/integrator.h)
class Integrator : public Object {
public:
    virtual bool render(Scene *scene, Sensor *sensor) = 0;
    MSK_DECLARE_CLASS()
protected:
    explicit Integrator(const Properties &) {}
};

// ---- Scene (reference include/misaki/render/scene.h, src/librender/scene.cpp:26-64)
class Scene : public Object {
public:
    explicit Scene(const Properties &props);
    const std::vector<ref<Shape>> &shapes() const { return m_shapes; }
    const std::vector<ref<Emitter>> &emitters() const { return m_emitters; }
    Emitter *environment() const { return m_environment.get(); }
    Sensor *sensor() const { return m_sensor.get(); }
    Integrator *integrator() const { return m_integrator.get(); }
    MSK_DECLARE_CLASS()
private:
    std::vector<ref<Shape>> m_shapes;
    std::vector<ref<Emitter>> m_emitters;
    ref<Emitter> m_environment;
    ref<Sensor> m_sensor;
    ref<Integrator> m_integrator;
};

// ---- flattening of the object graph into the C ABI's POD description (include/misaki_b200.h)
class GpuSceneBuilder {
public:
    explicit GpuSceneBuilder(const Scene *scene);
    const MskSceneDesc &desc() const { return m_desc; }
    int add_spectrum(const MskSpectrum &s, const float *table = nullptr, size_t table_size = 0);
    int add_bsdf(const MskBsdf &b);
    // memoised by object identity so shared (referenced) plugins are described once
    int spectrum_id(const Texture *t);
    int bsdf_id(const BSDF *b);
    int medium_id(const Medium *m); // -1 for nullptr
    bool within_conductor = false; // conductor eta / k: unbounded spectra (builder decision)

private:
    MskSceneDesc m_desc{};
    std::vector<MskMesh> m_meshes;
    std::vector<MskBsdf> m_bsdfs;
    std::vector<MskEmitter> m_emitters;
    std::vector<MskSpectrum> m_spectra;
    std::vector<float> m_tables;
    std::vector<MskMedium> m_media;
    std::map<const void *, int> m_spectrum_ids, m_bsdf_ids, m_medium_ids;
    void finish();
};

// HDRFilm::image: XYZAW -> RGBA (hdrfilm.cpp:48-90); image writers (OpenEXR scanline, uncompressed; PFM)
void develop_xyzaw(const float *film, size_t npixels, float *rgba);
void develop_channels(const float *film, size_t npixels, size_t nchannels, float *out); // with AOV channels (/ W)
void write_exr_rgba(const std::string &filename, const float *rgba, uint32_t width, uint32_t height);
void write_exr_channels(const std::string &filename, const std::vector<std::string> &names, const float *pixels, uint32_t width,
                        uint32_t height);
void write_pfm_rgb(const std::string &filename, const float *rgba, uint32_t width, uint32_t height);

// rgb2spec (reference ext/rgb2spec/rgb2spec.c:12-47,77-119 through src/librender/srgb.cpp:11-30)
Color3 srgb_model_fetch(const Color3 &c);

} // namespace misaki
