// The plugins a misaki scene file can name (the MSK_REGISTER_INSTANCE list of the reference build, SURVEY.md
// section 8b): obj | d65 regular srgb srgb_d65 uniform | diffuse conductor roughconductor roughdielectric
// dielectric twosided | area constant | perspective | independent | hdrfilm | gaussian.
// Integrators live in gpu_integrator.cpp.  Each constructor cites the reference constructor it restates.
#include "render.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

#include "../csrc/spectral_tables.h"

namespace misaki {

// =========================================================================================== base classes
MSK_IMPLEMENT_CLASS(Texture, Object, "texture")
MSK_IMPLEMENT_CLASS(BSDF, Object, "bsdf")
MSK_IMPLEMENT_CLASS(Emitter, Object, "emitter")
MSK_IMPLEMENT_CLASS(PhaseFunction, Object, "phase")
MSK_IMPLEMENT_CLASS(Medium, Object, "medium")
MSK_IMPLEMENT_CLASS(Shape, Object, "shape")
MSK_IMPLEMENT_CLASS(Mesh, Shape)
MSK_IMPLEMENT_CLASS(ReconstructionFilter, Object, "rfilter")
MSK_IMPLEMENT_CLASS(ImageBlock, Object)
MSK_IMPLEMENT_CLASS(Film, Object, "film")
MSK_IMPLEMENT_CLASS(Sampler, Object, "sampler")
MSK_IMPLEMENT_CLASS(Sensor, Object, "sensor")
MSK_IMPLEMENT_CLASS(Integrator, Object, "integrator")

ref<Texture> Texture::D65(float scale) { // texture.cpp:26-37
    Properties p("d65");
    p.set_float("scale", scale);
    ref<Texture> t = InstanceManager::get()->create_instance<Texture>(p);
    auto expanded = t->expand();
    return ref<Texture>(static_cast<Texture *>(expanded.at(0).get()));
}

Emitter::Emitter(const Properties &props) { // emitter.cpp:7-15
    m_world_transform = props.transform("to_world", Transform4f());
}

Shape::Shape(const Properties &props) : m_id(props.id()) { // shape.cpp:14-48
    m_world_transform = props.transform("to_world", Transform4f());
    for (auto &[name, obj] : props.objects()) {
        auto *emitter = dynamic_cast<Emitter *>(obj.get());
        auto *bsdf = dynamic_cast<BSDF *>(obj.get());
        auto *medium = dynamic_cast<Medium *>(obj.get());
        if (emitter) {
            if (m_emitter) Throw("Only one light can be specified by a shape.");
            m_emitter = emitter;
        } else if (bsdf) {
            if (m_bsdf) Throw("Only one bsdf can be specified by a shape.");
            m_bsdf = bsdf;
        } else if (medium) { // shape.cpp:28-40: children named "interior" / "exterior"; other names are dropped
            if (name == "interior") {
                if (m_interior_medium) Throw("Only a single interior medium can be specified per shape.");
                m_interior_medium = medium;
            } else if (name == "exterior") {
                if (m_exterior_medium) Throw("Only a single exterior medium can be specified per shape.");
                m_exterior_medium = medium;
            }
        } else {
            Throw("Tired to add unsuppored object of type \"%s\"", obj->to_string().c_str());
        }
    }
    if (!m_bsdf) m_bsdf = InstanceManager::get()->create_instance<BSDF>(Properties("diffuse"));
}
void Shape::set_children() { if (m_emitter) m_emitter->set_shape(this); }

Medium::Medium(const Properties &props) : m_id(props.id()) { // medium.cpp:13-31
    for (auto &[name, obj] : props.objects()) {
        auto *phase = dynamic_cast<PhaseFunction *>(obj.get());
        if (phase) {
            if (m_phase_function) Throw("Only a single phase function can be specified per medium");
            m_phase_function = phase;
        }
    }
    if (!m_phase_function) m_phase_function = InstanceManager::get()->create_instance<PhaseFunction>(Properties("isotropic"));
}

Mesh::Mesh(const Properties &props) : Shape(props) { m_to_world = props.transform("to_world", Transform4f()); } // mesh.cpp:11-13

Film::Film(const Properties &props) { // film.cpp:9-39
    m_width = (uint32_t) props.int_("width", 640);
    m_height = (uint32_t) props.int_("height", 320);
    for (auto &[name, obj] : props.objects()) {
        auto *rf = dynamic_cast<ReconstructionFilter *>(obj.get());
        if (rf) {
            if (m_filter) Throw("A film can only have one reconstruction filter.");
            m_filter = rf;
        }
    }
    if (!m_filter) m_filter = InstanceManager::get()->create_instance<ReconstructionFilter>(Properties("gaussian"));
}

Sampler::Sampler(const Properties &props) { // sampler.cpp:7-10
    m_sample_count = (uint32_t) props.int_("sample_count", 1);
    m_base_seed = (uint64_t) props.int_("base_seed", 0);
}

Sensor::Sensor(const Properties &props) { // sensor.cpp:9-44
    m_world_transform = props.transform("to_world", Transform4f());
    for (auto &[name, obj] : props.objects()) {
        auto *film = dynamic_cast<Film *>(obj.get());
        auto *sampler = dynamic_cast<Sampler *>(obj.get());
        auto *medium = dynamic_cast<Medium *>(obj.get());
        if (medium) { // sensor.cpp:12-18
            if (m_medium) Throw("Only a single medium can be specified per endpoint");
            m_medium = medium;
        } else if (film) {
            if (m_film) Throw("Camera can only have one film.");
            m_film = film;
        } else if (sampler) {
            if (m_sampler) Throw("Can only have one samplelr.");
            m_sampler = sampler;
        }
    }
    // the reference defaults to "rgbfilm" (sensor.cpp:36-37), which its build no longer compiles; hdrfilm is the
    // one film that exists, so a sensor without a film gets that
    if (!m_film) m_film = InstanceManager::get()->create_instance<Film>(Properties("hdrfilm"));
    if (!m_sampler) m_sampler = InstanceManager::get()->create_instance<Sampler>(Properties("independent"));
    m_aspect = (float) m_film->width() / (float) m_film->height();
}

void ReconstructionFilter::init_discretization() { // rfilter.cpp:12-27
    m_values.assign(MSK_FILTER_RESOLUTION + 1, 0.f);
    float sum = 0.f;
    for (int i = 0; i < MSK_FILTER_RESOLUTION; ++i) {
        float v = eval((m_radius * i) / MSK_FILTER_RESOLUTION);
        m_values[i] = v;
        sum += v;
    }
    m_values[MSK_FILTER_RESOLUTION] = 0.f;
    m_scale_factor = MSK_FILTER_RESOLUTION / m_radius;
    m_border_size = (uint32_t) std::ceil(m_radius - .5f);
    sum *= 2 * m_radius / MSK_FILTER_RESOLUTION;
    float normalization = 1.f / sum;
    for (int i = 0; i < MSK_FILTER_RESOLUTION; ++i) m_values[i] *= normalization;
}

// =========================================================================================== spectra
namespace {

MskSpectrum blank_spectrum(MskSpectrumKind kind) {
    MskSpectrum s{};
    s.kind = kind;
    return s;
}

class UniformSpectrum final : public Texture { // spectra/uniform.cpp:14-26
public:
    explicit UniformSpectrum(const Properties &props) : Texture(props) { m_value = props.float_("value"); }
    int describe(GpuSceneBuilder &b) const override {
        MskSpectrum s = blank_spectrum(MSK_SPEC_UNIFORM);
        s.value = m_value;
        return b.add_spectrum(s);
    }
    float mean() const override { return m_value; }
    float value() const { return m_value; }
    MSK_DECLARE_CLASS()
private:
    float m_value;
};

class RegularSpectrum final : public Texture { // spectra/regular.cpp:114-127
public:
    explicit RegularSpectrum(const Properties &props) : Texture(props) {
        m_lambda_min = props.float_("lambda_min");
        m_lambda_max = props.float_("lambda_max");
        size_t size = (size_t) props.int_("size");
        const float *values = (const float *) props.pointer("values");
        if (size < 2) Throw("ContinuousDistribution: needs at least two entries!");
        if (!(m_lambda_min < m_lambda_max)) Throw("ContinuousDistribution: invalid range!");
        m_values.assign(values, values + size);
        // SpectrumContinuousDistribution::update (regular.cpp:31-58): trapezoid masses in double; negative entries and
        // an all-zero table are errors (so a black <rgb> inside an emitter throws, as in the reference)
        double interval = (double(m_lambda_max) - double(m_lambda_min)) / double(size - 1);
        bool mass = false;
        for (size_t i = 0; i + 1 < size; ++i) {
            double y0 = m_values[i], y1 = m_values[i + 1];
            if (y0 < 0. || y1 < 0.) Throw("ContinuousDistribution: entries must be non-negative!");
            mass |= 0.5 * interval * (y0 + y1) > 0.;
        }
        if (!mass) Throw("ContinuousDistribution: no probability mass found!");
    }
    int describe(GpuSceneBuilder &b) const override {
        MskSpectrum s = blank_spectrum(MSK_SPEC_REGULAR);
        s.lambda_min = m_lambda_min; s.lambda_max = m_lambda_max;
        return b.add_spectrum(s, m_values.data(), m_values.size());
    }
    const std::vector<float> &values() const { return m_values; }
    MSK_DECLARE_CLASS()
private:
    float m_lambda_min, m_lambda_max;
    std::vector<float> m_values;
};

class D65Spectrum final : public Texture { // spectra/d65.cpp:29-50
public:
    explicit D65Spectrum(const Properties &props) : Texture(props) {
        m_scale = props.float_("scale", 1.f);
        m_scale *= 1.f / 10568.f;
    }
    std::vector<ref<Object>> expand() const override {
        Properties p("regular");
        p.set_float("lambda_min", 360);
        p.set_float("lambda_max", 830);
        p.set_int("size", 95);
        float tmp[95];
        for (size_t i = 0; i < 95; ++i) tmp[i] = msk_cie_d65_rows[i][3] * m_scale;
        p.set_pointer("values", (const void *) &tmp[0]);
        return { ref<Object>(InstanceManager::get()->create_instance<Texture>(p).get()) };
    }
    int describe(GpuSceneBuilder &b) const override {
        auto e = expand();
        return static_cast<const Texture *>(e[0].get())->describe(b);
    }
    MSK_DECLARE_CLASS()
private:
    float m_scale;
};

class SRGBReflectanceSpectrum final : public Texture { // spectra/srgb.cpp:14-23
public:
    explicit SRGBReflectanceSpectrum(const Properties &props) : Texture(props) {
        m_color = props.color("color");
        m_value = srgb_model_fetch(m_color);
    }
    int describe(GpuSceneBuilder &b) const override {
        if (b.within_conductor) {
            // conductor eta / k given as <rgb>: values exceed 1, which rgb2spec_fetch would clamp (rgb2spec.c:81-82)
            // and the reference's eval_3 path is unimplemented.  Builder decision shared with the
            // oracle: scale = 2 max(rgb), coefficients of rgb / scale, value = scale * srgb_model_eval.
            Color3 c = m_color;
            float scale = std::max(c.r, std::max(c.g, c.b)) * 2.f;
            if (scale != 0.f) { c.r /= scale; c.g /= scale; c.b /= scale; }
            Color3 v = srgb_model_fetch(c);
            MskSpectrum s = blank_spectrum(MSK_SPEC_SRGB_UNBOUNDED);
            s.c[0] = v.r; s.c[1] = v.g; s.c[2] = v.b; s.value = scale;
            return b.add_spectrum(s);
        }
        MskSpectrum s = blank_spectrum(MSK_SPEC_SRGB);
        s.c[0] = m_value.r; s.c[1] = m_value.g; s.c[2] = m_value.b;
        return b.add_spectrum(s);
    }
    MSK_DECLARE_CLASS()
private:
    Color3 m_color, m_value;
};

class SRGBEmitterSpectrum final : public Texture { // spectra/srgb_d65.cpp:14-36
public:
    explicit SRGBEmitterSpectrum(const Properties &props) : Texture(props) {
        Color3 color = props.color("color");
        float scale = std::max(color.r, std::max(color.g, color.b)) * 2.f;
        if (scale != 0.f) { color.r /= scale; color.g /= scale; color.b /= scale; }
        m_value = srgb_model_fetch(color);
        Properties p2("d65");
        p2.set_float("scale", props.float_("scale", 1.f) * scale);
        ref<Texture> d65 = InstanceManager::get()->create_instance<Texture>(p2);
        m_d65 = static_cast<Texture *>(d65->expand().at(0).get());
    }
    int describe(GpuSceneBuilder &b) const override {
        auto *reg = static_cast<const RegularSpectrum *>(m_d65.get());
        MskSpectrum s = blank_spectrum(MSK_SPEC_SRGB_D65);
        s.c[0] = m_value.r; s.c[1] = m_value.g; s.c[2] = m_value.b;
        s.lambda_min = 360.f; s.lambda_max = 830.f;
        return b.add_spectrum(s, reg->values().data(), reg->values().size());
    }
    MSK_DECLARE_CLASS()
private:
    Color3 m_value;
    ref<Texture> m_d65;
};

class CheckerboardTexture final : public Texture { // textures/checkerboard.cpp:8-46
public:
    explicit CheckerboardTexture(const Properties &props) : Texture(props) {
        m_color0 = props.texture("color0", .4f);
        m_color1 = props.texture("color1", .2f);
        m_to_uv = props.transform("to_uv", Transform4f());
    }
    int describe(GpuSceneBuilder &b) const override {
        MskSpectrum s = blank_spectrum(MSK_SPEC_CHECKERBOARD);
        s.child0 = m_color0->describe(b); // children first: their ids are smaller than the checkerboard's
        s.child1 = m_color1->describe(b);
        for (int r = 0; r < 2; ++r) // Transform4f::extract(): the top-left 3x3 (transform.h:142-148)
            for (int c = 0; c < 3; ++c) s.to_uv[r * 3 + c] = m_to_uv.m[r * 4 + c];
        return b.add_spectrum(s);
    }
    float mean() const override { return m_color0->mean() + m_color1->mean(); } // (sic) checkerboard.cpp:44
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_color0, m_color1;
    Transform4f m_to_uv;
};

MSK_IMPLEMENT_PLUGIN(UniformSpectrum, Texture, "uniform")
MSK_IMPLEMENT_PLUGIN(CheckerboardTexture, Texture, "checkerboard")
MSK_IMPLEMENT_PLUGIN(RegularSpectrum, Texture, "regular")
MSK_IMPLEMENT_PLUGIN(D65Spectrum, Texture, "d65")
MSK_IMPLEMENT_PLUGIN(SRGBReflectanceSpectrum, Texture, "srgb")
MSK_IMPLEMENT_PLUGIN(SRGBEmitterSpectrum, Texture, "srgb_d65")

// scalar parameters that reach a BSDF as textures (<float name="alpha" .../> -> "uniform", properties.cpp:206-211)
float scalar_of(const ref<Texture> &t, const char *what) {
    auto *u = dynamic_cast<const UniformSpectrum *>(t.get());
    if (!u) Throw("\"%s\" must be a scalar (<float>) in this build: spectrally varying roughness is undefined in the reference "
                  "(Texture::eval_1 is unimplemented, texture.cpp:14-24)", what);
    return u->value();
}

// =========================================================================================== BSDFs
MskBsdf blank_bsdf(MskBsdfType type) {
    MskBsdf b{};
    b.type = type;
    b.reflectance = b.transmittance = b.eta = b.k = -1;
    b.alpha_u = b.alpha_v = 0.1f;
    b.int_ior = 1.5046f; b.ext_ior = 1.00028f;
    b.distribution = 1;
    return b;
}

class SmoothDiffuse final : public BSDF { // bsdfs/diffuse.cpp:11-17
public:
    explicit SmoothDiffuse(const Properties &props) : BSDF(props) { m_reflectance = props.texture("reflectance", 0.5f); }
    int describe(GpuSceneBuilder &b) const override {
        MskBsdf d = blank_bsdf(MSK_BSDF_DIFFUSE);
        d.reflectance = b.spectrum_id(m_reflectance.get());
        return b.add_bsdf(d);
    }
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_reflectance;
};

class ConductorBSDF final : public BSDF { // bsdfs/conductor.cpp:10-18 (stale RGB API in the reference; restated spectrally)
public:
    explicit ConductorBSDF(const Properties &props) : BSDF(props) {
        m_specular_reflectance = props.texture("specular_reflectance", 1.f);
        m_eta = props.texture("eta", 0.f);
        m_k = props.texture("k", 1.f);
    }
    int describe(GpuSceneBuilder &b) const override {
        MskBsdf d = blank_bsdf(MSK_BSDF_CONDUCTOR);
        d.reflectance = b.spectrum_id(m_specular_reflectance.get());
        b.within_conductor = true;
        d.eta = m_eta->describe(b); d.k = m_k->describe(b);
        b.within_conductor = false;
        return b.add_bsdf(d);
    }
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_specular_reflectance, m_eta, m_k;
};

struct MicrofacetParams { // the parameter block shared by roughconductor.cpp:19-45 and roughdielectric.cpp:25-51
    int distribution = 0; // 0 beckmann (reference default), 1 ggx
    bool sample_visible = false;
    float alpha_u = 0.1f, alpha_v = 0.1f;
    void read(const Properties &props, bool lower) {
        if (props.has_property("distribution")) {
            std::string distr = props.string("distribution");
            if (lower) distr = string::to_lower(distr);
            if (distr == "beckmann") distribution = 0;
            else if (distr == "ggx") distribution = 1;
            else Throw("Specified an invalid distribution \"%s\", must be \"beckmann\" or \"ggx\"!", distr.c_str());
        }
        sample_visible = props.bool_("sample_visible", false);
        if (props.has_property("alpha_u") || props.has_property("alpha_v")) {
            if (!props.has_property("alpha_u") || !props.has_property("alpha_v"))
                Throw("Microfacet model: both 'alpha_u' and 'alpha_v' must be specified.");
            if (props.has_property("alpha")) Throw("Microfacet model: please specify either 'alpha' or 'alpha_u'/'alpha_v'.");
            alpha_u = scalar_of(props.texture("alpha_u"), "alpha_u");
            alpha_v = scalar_of(props.texture("alpha_v"), "alpha_v");
        } else if (props.has_property("alpha")) {
            alpha_u = alpha_v = scalar_of(props.texture("alpha"), "alpha");
        }
        if (distribution == 0)
            Throw("the \"beckmann\" microfacet distribution is a stub in the reference (eval returns 0 and sample has no return "
                  "value, microfacet.h:113-115,134-136); specify <string name=\"distribution\" value=\"ggx\"/>");
    }
};

class RoughConductor final : public BSDF { // bsdfs/roughconductor.cpp:14-50
public:
    explicit RoughConductor(const Properties &props) : BSDF(props) {
        if (props.has_property("eta")) {
            m_eta = props.texture("eta", 0.f);
            m_k = props.texture("k", 1.f);
        } else {
            // the reference leaves m_eta null and dereferences it in sample() (roughconductor.cpp:15-18,78)
            Throw("roughconductor: \"eta\" (and \"k\") must be specified");
        }
        m_mf.read(props, false);
        m_specular_reflectance = props.texture("specular_reflectance", 1.f);
    }
    int describe(GpuSceneBuilder &b) const override {
        MskBsdf d = blank_bsdf(MSK_BSDF_ROUGHCONDUCTOR);
        d.reflectance = b.spectrum_id(m_specular_reflectance.get());
        b.within_conductor = true;
        d.eta = m_eta->describe(b); d.k = m_k->describe(b);
        b.within_conductor = false;
        d.alpha_u = m_mf.alpha_u; d.alpha_v = m_mf.alpha_v; d.distribution = m_mf.distribution; d.sample_visible = m_mf.sample_visible;
        return b.add_bsdf(d);
    }
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_specular_reflectance, m_eta, m_k;
    MicrofacetParams m_mf;
};

class RoughDielectric final : public BSDF { // bsdfs/roughdielectric.cpp:15-55
public:
    explicit RoughDielectric(const Properties &props) : BSDF(props) {
        m_specular_reflectance = props.texture("specular_reflectance", 1.f);
        m_specular_transmittance = props.texture("specular_transmittance", 1.f);
        m_int_ior = props.float_("int_ior", 1.5046f);
        m_ext_ior = props.float_("ext_ior", 1.00028f);
        if (m_int_ior < 0.f || m_ext_ior < 0.f || m_int_ior == m_ext_ior)
            Throw("The interior and exterior indices of refraction must be positive and differ!");
        m_mf.read(props, true);
    }
    bool has_transmission() const override { return true; }
    int describe(GpuSceneBuilder &b) const override {
        MskBsdf d = blank_bsdf(MSK_BSDF_ROUGHDIELECTRIC);
        d.reflectance = b.spectrum_id(m_specular_reflectance.get());
        d.transmittance = b.spectrum_id(m_specular_transmittance.get());
        d.int_ior = m_int_ior; d.ext_ior = m_ext_ior;
        d.alpha_u = m_mf.alpha_u; d.alpha_v = m_mf.alpha_v; d.distribution = m_mf.distribution; d.sample_visible = m_mf.sample_visible;
        return b.add_bsdf(d);
    }
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_specular_reflectance, m_specular_transmittance;
    float m_int_ior, m_ext_ior;
    MicrofacetParams m_mf;
};

class SmoothDielectric final : public BSDF { // bsdfs/dielectric.cpp:12-24
public:
    explicit SmoothDielectric(const Properties &props) : BSDF(props) {
        m_int_ior = props.float_("int_ior", 1.49f);
        m_ext_ior = props.float_("ext_ior", 1.00028f);
        if (m_int_ior < 0.f || m_ext_ior < 0.f || m_int_ior == m_ext_ior)
            Throw("The interior and exterior indices of refraction must be positive and differ!");
        m_specular_reflectance = props.texture("specular_reflectance", 1.f);
        m_specular_transmittance = props.texture("specular_transmittance", 1.f);
    }
    bool has_transmission() const override { return true; }
    int describe(GpuSceneBuilder &b) const override {
        MskBsdf d = blank_bsdf(MSK_BSDF_DIELECTRIC);
        d.reflectance = b.spectrum_id(m_specular_reflectance.get());
        d.transmittance = b.spectrum_id(m_specular_transmittance.get());
        d.int_ior = m_int_ior; d.ext_ior = m_ext_ior;
        return b.add_bsdf(d);
    }
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_specular_reflectance, m_specular_transmittance;
    float m_int_ior, m_ext_ior;
};

class TwoSidedBRDF final : public BSDF { // bsdfs/twosided.cpp:11-36
public:
    explicit TwoSidedBRDF(const Properties &props) : BSDF(props) {
        auto bsdfs = props.objects();
        if (!bsdfs.empty()) m_brdf[0] = dynamic_cast<BSDF *>(bsdfs[0].second.get());
        if (bsdfs.size() == 2) m_brdf[1] = dynamic_cast<BSDF *>(bsdfs[1].second.get());
        else if (bsdfs.size() > 2) Throw("At most two nested BSDFs can be specified!");
        if (!m_brdf[0]) Throw("A nested one-sided material is required!");
        if (!m_brdf[1]) m_brdf[1] = m_brdf[0];
        if (m_brdf[0]->has_transmission() || m_brdf[1]->has_transmission())
            Throw("Only materials without a transmission component can be nested!");
        if (m_brdf[0].get() != m_brdf[1].get())
            Throw("twosided: two different nested BRDFs are not supported by this build (one BRDF on both sides is)");
    }
    int describe(GpuSceneBuilder &b) const override {
        // a private entry for the nested BRDF (described directly, not through the memo), which
        // GpuSceneBuilder::bsdf_id marks two-sided: a negative return value -(entry + 2) requests that
        return -(m_brdf[0]->describe(b) + 2);
    }
    MSK_DECLARE_CLASS()
private:
    ref<BSDF> m_brdf[2];
};

MSK_IMPLEMENT_PLUGIN(SmoothDiffuse, BSDF, "diffuse")
MSK_IMPLEMENT_PLUGIN(ConductorBSDF, BSDF, "conductor")
MSK_IMPLEMENT_PLUGIN(RoughConductor, BSDF, "roughconductor")
MSK_IMPLEMENT_PLUGIN(RoughDielectric, BSDF, "roughdielectric")
MSK_IMPLEMENT_PLUGIN(SmoothDielectric, BSDF, "dielectric")
MSK_IMPLEMENT_PLUGIN(TwoSidedBRDF, BSDF, "twosided")

// =========================================================================================== emitters
class AreaLight final : public Emitter { // emitters/area.cpp:12-16
public:
    explicit AreaLight(const Properties &props) : Emitter(props) { m_radiance = props.texture("radiance", Texture::D65(1.f)); }
    bool is_surface() const override { return true; }
    MskEmitterType gpu_type() const override { return MSK_EMITTER_AREA; }
    MSK_DECLARE_CLASS()
};

class ConstantBackgroundEmitter final : public Emitter { // emitters/constant.cpp:14-19
public:
    explicit ConstantBackgroundEmitter(const Properties &props) : Emitter(props) { m_radiance = props.texture("radiance", Texture::D65(1.f)); }
    bool is_environment() const override { return true; }
    MskEmitterType gpu_type() const override { return MSK_EMITTER_CONSTANT; }
    MSK_DECLARE_CLASS()
};
MSK_IMPLEMENT_PLUGIN(AreaLight, Emitter, "area")
MSK_IMPLEMENT_PLUGIN(ConstantBackgroundEmitter, Emitter, "constant")

// =========================================================================================== media
class IsotropicPhaseFunction final : public PhaseFunction { // phase/isotropic.cpp:9-36
public:
    explicit IsotropicPhaseFunction(const Properties &props) : PhaseFunction(props) {}
    MskPhaseType gpu_type() const override { return MSK_PHASE_ISOTROPIC; }
    MSK_DECLARE_CLASS()
};
MSK_IMPLEMENT_PLUGIN(IsotropicPhaseFunction, PhaseFunction, "isotropic")

// media/homogeneous.cpp:10-19.  The reference reads sigma_a / sigma_s with props.color() (an RGB triple of the stale
// RGB pipeline, default 1); on the spectral pipeline they are textures like every other colour parameter: <rgb>
// gives an UNBOUNDED upsampled spectrum (coefficients, not reflectances -- the conductor eta / k rule), <float> /
// <spectrum> a uniform / tabulated one.  "scale" is read and stored but never applied, as in the reference.
class HomogeneousMedium final : public Medium {
public:
    explicit HomogeneousMedium(const Properties &props) : Medium(props) {
        m_sigma_a = props.texture("sigma_a", 1.f);
        m_sigma_s = props.texture("sigma_s", 1.f);
        m_scale = props.float_("scale", 1.f);
    }
    void describe(GpuSceneBuilder &b, MskMedium &out) const override {
        bool saved = b.within_conductor;
        b.within_conductor = true; // unbounded spectra
        out.sigma_a = m_sigma_a->describe(b);
        out.sigma_s = m_sigma_s->describe(b);
        b.within_conductor = saved;
        out.phase = m_phase_function->gpu_type();
        out.scale = m_scale;
    }
    MSK_DECLARE_CLASS()
private:
    ref<Texture> m_sigma_a, m_sigma_s;
    float m_scale;
};
MSK_IMPLEMENT_PLUGIN(HomogeneousMedium, Medium, "homogeneous")

// =========================================================================================== obj shape
static int to_uint(const std::string &str) { // shapes/obj.cpp:11-17; a leading '-' (relative index) is kept signed
    char *end_ptr = nullptr;
    long result = strtol(str.c_str(), &end_ptr, 10);
    if (*end_ptr != '\0') Throw("Could not parse integer value \"%s\"", str.c_str());
    return (int) result;
}
struct OBJVertex { // shapes/obj.cpp:19-38
    int p = -1, n = -1, uv = -1; // 1-based; -1 = absent (after make_absolute)
    OBJVertex() = default;
    // OBJ relative indices: -k names the k-th most recent element at the time of the face line
    void make_absolute(size_t np, size_t nuv, size_t nn, bool has_uv, bool has_n) {
        if (p < 0) p = (int) np + 1 + p;
        if (has_uv && uv < 0) uv = (int) nuv + 1 + uv;
        if (has_n && n < 0) n = (int) nn + 1 + n;
    }
    bool has_uv_ = false, has_n_ = false;
    explicit OBJVertex(const std::string &s) {
        auto tokens = string::tokenize(s, "/", true);
        if (tokens.size() < 1 || tokens.size() > 3) Throw("Invalid vertex data: \"%s\"", s.c_str());
        p = to_uint(tokens[0]);
        if (tokens.size() >= 2 && !tokens[1].empty()) { uv = to_uint(tokens[1]); has_uv_ = true; }
        if (tokens.size() >= 3 && !tokens[2].empty()) { n = to_uint(tokens[2]); has_n_ = true; }
    }
    bool operator==(const OBJVertex &v) const { return v.p == p && v.n == n && v.uv == uv; }
};
struct OBJVertexHash {
    size_t operator()(const OBJVertex &v) const {
        size_t h = std::hash<int>()(v.p);
        h = h * 37 + std::hash<int>()(v.uv);
        h = h * 37 + std::hash<int>()(v.n);
        return h;
    }
};

class OBJMesh final : public Mesh { // shapes/obj.cpp:58-181
public:
    explicit OBJMesh(const Properties &props) : Mesh(props) {
        bool flip_tex_coords = props.bool_("filp_tex_coords", true); // (sic) the reference's parameter name
        std::string file_path = get_file_resolver()->resolve(props.string("filename"));
        size_t slash = file_path.find_last_of('/');
        m_name = slash == std::string::npos ? file_path : file_path.substr(slash + 1);
        Log(Info, "Loading mesh from \"%s\"", m_name.c_str());
        std::ifstream is(file_path);
        if (!is) Throw("Error while loading OBJ file \"%s\": file not found", m_name.c_str());
        std::vector<Vector3f> vertices, normals;
        std::vector<std::array<float, 2>> texcoords;
        std::vector<uint32_t> triangles;
        std::vector<OBJVertex> obj_vertices;
        std::unordered_map<OBJVertex, uint32_t, OBJVertexHash> vertex_map;
        std::string line_str;
        while (std::getline(is, line_str)) {
            std::istringstream line(line_str);
            std::string prefix;
            line >> prefix;
            if (prefix == "v") {
                Vector3f p;
                line >> p.x >> p.y >> p.z;
                vertices.push_back(m_to_world.apply_point(p));
            } else if (prefix == "vt") {
                std::array<float, 2> tc{};
                line >> tc[0] >> tc[1];
                if (flip_tex_coords) tc[1] = 1.f - tc[1];
                texcoords.push_back(tc);
            } else if (prefix == "vn") {
                Vector3f n;
                line >> n.x >> n.y >> n.z;
                n = m_to_world.apply_normal(n);
                float len = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
                if (len > 0.f) { n.x /= len; n.y /= len; n.z /= len; }
                normals.push_back(n);
            } else if (prefix == "f") {
                // triangle, quad -> (v1 v2 v3) (v4 v1 v3) as obj.cpp:104-118, and the same fan for n-gons:
                // (v1 v2 v3) (v4 v1 v3) (v5 v1 v4) ...  (the reference reads at most four corners)
                std::vector<OBJVertex> corners;
                std::string tok;
                while (line >> tok) {
                    OBJVertex c(tok);
                    c.make_absolute(vertices.size(), texcoords.size(), normals.size(), c.has_uv_, c.has_n_);
                    c.has_uv_ = c.has_n_ = false;
                    corners.push_back(c);
                }
                if (corners.size() < 3) Throw("Error while loading OBJ file \"%s\": face with fewer than 3 vertices", m_name.c_str());
                std::vector<OBJVertex> verts = { corners[0], corners[1], corners[2] };
                for (size_t j = 3; j < corners.size(); ++j) { verts.push_back(corners[j]); verts.push_back(corners[0]); verts.push_back(corners[j - 1]); }
                int n_vertices = (int) verts.size();
                for (int i = 0; i < n_vertices; ++i) {
                    auto it = vertex_map.find(verts[i]);
                    if (it == vertex_map.end()) {
                        vertex_map[verts[i]] = (uint32_t) obj_vertices.size();
                        triangles.push_back((uint32_t) obj_vertices.size());
                        obj_vertices.push_back(verts[i]);
                    } else {
                        triangles.push_back(it->second);
                    }
                }
            }
        }
        m_vertex_count = (uint32_t) obj_vertices.size();
        m_face_count = (uint32_t) (triangles.size() / 3);
        m_normal_offset = normals.empty() ? 0 : 3;
        m_texcoord_offset = texcoords.empty() ? 0 : 6;
        m_faces.assign(triangles.begin(), triangles.begin() + (size_t) m_face_count * 3);
        m_vertices.assign((size_t) m_vertex_count * 8, 0.f); // the reference leaves absent attributes uninitialised; zero here
        for (size_t i = 0; i < obj_vertices.size(); ++i) {
            const OBJVertex &v = obj_vertices[i];
            if (v.p < 1 || (size_t) v.p > vertices.size()) Throw("Error while loading OBJ file \"%s\": vertex index %d out of range", m_name.c_str(), v.p);
            float *o = &m_vertices[i * 8];
            o[0] = vertices[v.p - 1].x; o[1] = vertices[v.p - 1].y; o[2] = vertices[v.p - 1].z;
            if (v.n != -1) {
                if (v.n < 1 || (size_t) v.n > normals.size()) Throw("Error while loading OBJ file \"%s\": normal index %d out of range", m_name.c_str(), v.n);
                o[3] = normals[v.n - 1].x; o[4] = normals[v.n - 1].y; o[5] = normals[v.n - 1].z;
            }
            if (v.uv != -1) {
                if (v.uv < 1 || (size_t) v.uv > texcoords.size()) Throw("Error while loading OBJ file \"%s\": texcoord index %d out of range", m_name.c_str(), v.uv);
                o[6] = texcoords[v.uv - 1][0]; o[7] = texcoords[v.uv - 1][1];
            }
        }
        Log(Info, "\"%s\": read %u faces, %u vertices", m_name.c_str(), m_face_count, m_vertex_count);
        set_children();
    }
    MSK_DECLARE_CLASS()
};
MSK_IMPLEMENT_PLUGIN(OBJMesh, Mesh, "obj")

// =========================================================================================== sensor / sampler / film / filter
class PerspectiveCamera final : public Sensor { // sensor.cpp:136-142 (ProjectiveCamera) + sensors/perspective.cpp:9-20
public:
    explicit PerspectiveCamera(const Properties &props) : Sensor(props) {
        m_near_clip = props.float_("near_clip", 1e-2f);
        m_far_clip = props.float_("far_clip", 1e4f);
        m_fov = props.float_("fov", 30);
        // camera_to_sample = scale(W,H,1) * scale(-1/2, -aspect/2, 1) * translate(-1, -1/aspect, 0) * perspective
        m_camera_to_sample = Transform4f::scale(Vector3f{ (float) m_film->width(), (float) m_film->height(), 1.f }) *
                             Transform4f::scale(Vector3f{ -0.5f, -0.5f * m_aspect, 1.f }) *
                             Transform4f::translate(Vector3f{ -1.f, -1.f / m_aspect, 0.f }) *
                             Transform4f::perspective(m_fov, m_near_clip, m_far_clip);
    }
    void describe(MskCamera &cam) const override {
        memcpy(cam.sample_to_camera, m_camera_to_sample.inv, sizeof(float) * 16);
        memcpy(cam.to_world, m_world_transform.m, sizeof(float) * 16);
        cam.near_clip = m_near_clip; cam.far_clip = m_far_clip;
        cam.width = m_film->width(); cam.height = m_film->height();
        const ReconstructionFilter *f = m_film->filter();
        if (f->values().size() != 33) Throw("reconstruction filter table must have 33 entries");
        cam.filter_radius = f->radius();
        memcpy(cam.filter_table, f->values().data(), sizeof(float) * 33);
    }
    MSK_DECLARE_CLASS()
private:
    float m_near_clip, m_far_clip, m_fov;
    Transform4f m_camera_to_sample;
};
MSK_IMPLEMENT_PLUGIN(PerspectiveCamera, Sensor, "perspective")

class IndependentSampler final : public Sampler { // samplers/independent.cpp:9-12
public:
    explicit IndependentSampler(const Properties &props) : Sampler(props) {}
    MSK_DECLARE_CLASS()
};
MSK_IMPLEMENT_PLUGIN(IndependentSampler, Sampler, "independent")

class GaussianFilter final : public ReconstructionFilter { // filters/gaussian.cpp:9-20
public:
    explicit GaussianFilter(const Properties &props) : ReconstructionFilter(props) {
        m_stddev = props.float_("stddev", 0.5f);
        m_radius = 4 * m_stddev;
        m_alpha = -1.f / (2.f * m_stddev * m_stddev);
        m_bias = std::exp(m_alpha * m_radius * m_radius);
        init_discretization();
    }
    float eval(float x) const override { return std::max(0.f, std::exp(m_alpha * x * x) - m_bias); }
    MSK_DECLARE_CLASS()
private:
    float m_stddev, m_alpha, m_bias;
};
MSK_IMPLEMENT_PLUGIN(GaussianFilter, ReconstructionFilter, "gaussian")

class HDRFilm final : public Film { // films/hdrfilm.cpp:16-112
public:
    explicit HDRFilm(const Properties &props) : Film(props) {
        std::string file_format = string::to_lower(props.string("file_format", "openexr"));
        std::string pixel_format = string::to_lower(props.string("pixel_format", "rgba"));
        m_dest_file = props.string("filename", "");
        if (file_format != "openexr" && file_format != "exr" && file_format != "pfm")
            Throw("The \"file_format\" parameter must either be equal to \"openexr\" or \"pfm\", found %s instead.", file_format.c_str());
        m_pfm = file_format == "pfm";
        if (pixel_format != "rgba" && pixel_format != "rgb")
            Throw("The \"pixel_format\" parameter must either be equal to \"rgb\" or \"rgba\". Found %s.", pixel_format.c_str());
    }
    void prepare(const std::vector<std::string> &channels) override { // hdrfilm.cpp:30-41
        for (size_t i = 1; i < channels.size(); ++i)
            if (channels[i] == channels[i - 1]) Throw("Film::prepare(): duplicate channel name \"%s\"", channels[i].c_str());
        m_channels = channels;
        m_storage.assign((size_t) m_width * m_height * channels.size(), 0.f);
    }
    void put(const ImageBlock *block) override { // hdrfilm.cpp:43-46: accumulate (here a border-less, film-sized block)
        if (block->width() != m_width || block->height() != m_height || block->channel_count() != m_channels.size())
            Throw("HDRFilm::put: block does not match the film");
        const std::vector<float> &src = block->data();
        for (size_t i = 0; i < m_storage.size(); ++i) m_storage[i] += src[i];
    }
    void set_destination_file(const std::string &filename) override {
        size_t dot = filename.find_last_of('.'), slash = filename.find_last_of('/');
        std::string stem = (dot != std::string::npos && (slash == std::string::npos || dot > slash)) ? filename.substr(0, dot) : filename;
        m_dest_file = stem + (m_pfm ? ".pfm" : ".exr");
    }
    void develop() override { // hdrfilm.cpp:92-112
        if (m_dest_file.empty()) Throw("Destination file not specified, cannot develop.");
        if (m_channels.size() < 5) Throw("HDRFilm::develop: expected the X, Y, Z, A, W channels");
        const size_t nch = m_channels.size();
        std::vector<std::string> names = { "R", "G", "B", "A" }; // hdrfilm.cpp:52-59: RGBA, then the AOV channels
        names.insert(names.end(), m_channels.begin() + 5, m_channels.end());
        std::vector<float> image((size_t) m_width * m_height * (nch - 1));
        develop_channels(m_storage.data(), (size_t) m_width * m_height, nch, image.data());
        Log(Info, "Developing \"%s\" ..", m_dest_file.c_str());
        if (!m_pfm) { write_exr_channels(m_dest_file, names, image.data(), m_width, m_height); return; }
        std::vector<float> rgba((size_t) m_width * m_height * 4); // PFM holds RGB only
        for (size_t i = 0; i < (size_t) m_width * m_height; ++i)
            for (int c = 0; c < 4; ++c) rgba[i * 4 + c] = image[i * (nch - 1) + c];
        write_pfm_rgb(m_dest_file, rgba.data(), m_width, m_height);
    }
    const std::vector<std::string> &channels() const { return m_channels; }
    const std::vector<float> &storage() const { return m_storage; }
    const std::string &destination() const { return m_dest_file; }
    MSK_DECLARE_CLASS()
private:
    std::string m_dest_file;
    bool m_pfm = false;
    std::vector<std::string> m_channels;
    std::vector<float> m_storage;
};
MSK_IMPLEMENT_PLUGIN(HDRFilm, Film, "hdrfilm")
// "rgbfilm" is the reference Sensor's default and what assets/cbox/scene.xml names, but films/rgbfilm.cpp is
// written against the previous API generation and is not compiled (CMakeLists.txt:118-124): such scenes fail to
// load in the reference.  Here the name is served by HDRFilm (same XYZAW accumulation, float output).
static struct RgbFilmAlias_ { RgbFilmAlias_() { InstanceManager::get()->register_instance("rgbfilm", HDRFilm::m_class); } } rgbfilm_alias_;

} // namespace

// =========================================================================================== Scene
Scene::Scene(const Properties &props) { // scene.cpp:26-64 (accel_init happens when the GPU scene is created)
    for (auto &[name, obj] : props.objects()) {
        auto *shape = dynamic_cast<Shape *>(obj.get());
        auto *sensor = dynamic_cast<Sensor *>(obj.get());
        auto *integrator = dynamic_cast<Integrator *>(obj.get());
        auto *emitter = dynamic_cast<Emitter *>(obj.get());
        if (shape) {
            if (shape->is_emitter()) m_emitters.emplace_back(shape->emitter());
            m_shapes.emplace_back(shape);
        } else if (emitter) {
            if (!emitter->is_surface()) m_emitters.emplace_back(emitter);
            if (emitter->is_environment()) {
                if (m_environment) Throw("Can only have one environment light");
                m_environment = emitter;
            }
        } else if (sensor) {
            if (m_sensor) Throw("Can only have one camera.");
            m_sensor = sensor;
        } else if (integrator) {
            if (m_integrator) Throw("Can only have one integrator.");
            m_integrator = integrator;
        }
    }
    if (!m_integrator) {
        Log(Warn, "No integrator found! Instantiating a path tracer..");
        m_integrator = InstanceManager::get()->create_instance<Integrator>(Properties("path"));
    }
}
Class *Scene::m_class = new Class("Scene", "Object", "scene", [](const Properties &p) -> Object * { return new Scene(p); });
const Class *Scene::clazz() const { return m_class; }
static struct SceneRegister_ { SceneRegister_() { InstanceManager::get()->register_instance("scene", Scene::m_class); } } scene_register_instance_;

// =========================================================================================== flattening
GpuSceneBuilder::GpuSceneBuilder(const Scene *scene) {
    if (!scene->sensor()) Throw("The scene has no sensor");
    // emitters in Scene::m_emitters order; shapes in Scene::m_shapes order (== geomID)
    std::map<const Emitter *, int> emitter_index;
    for (auto &e : scene->emitters()) {
        MskEmitter d{};
        d.type = e->gpu_type();
        d.radiance = spectrum_id(e->radiance());
        d.shape = -1;
        emitter_index[e.get()] = (int) m_emitters.size();
        if (e->is_environment()) m_desc.environment = (int) m_emitters.size();
        m_emitters.push_back(d);
    }
    if (!scene->environment()) m_desc.environment = -1;
    for (auto &s : scene->shapes()) {
        auto *mesh = dynamic_cast<const Mesh *>(s.get());
        if (!mesh) Throw("Only triangle meshes are supported (got %s)", s->to_string().c_str());
        MskMesh m{};
        m.verts = mesh->vertices(); m.tris = mesh->faces();
        m.nverts = mesh->vertex_count(); m.ntris = mesh->face_count();
        m.bsdf = bsdf_id(mesh->bsdf());
        m.emitter = -1;
        if (mesh->is_emitter()) {
            m.emitter = emitter_index.at(mesh->emitter());
            m_emitters[m.emitter].shape = (int) m_meshes.size();
        }
        m.has_normals = mesh->has_vertex_normals(); m.has_uvs = mesh->has_vertex_texcoords();
        m.interior_medium = medium_id(mesh->interior_medium());
        m.exterior_medium = medium_id(mesh->exterior_medium());
        m_meshes.push_back(m);
    }
    m_desc.sensor_medium = medium_id(scene->sensor()->medium());
    scene->sensor()->describe(m_desc.camera);
    finish();
}
int GpuSceneBuilder::add_spectrum(const MskSpectrum &s, const float *table, size_t table_size) {
    MskSpectrum c = s;
    if (table) {
        c.table_offset = (uint32_t) m_tables.size();
        c.table_size = (uint32_t) table_size;
        m_tables.insert(m_tables.end(), table, table + table_size);
    }
    m_spectra.push_back(c);
    return (int) m_spectra.size() - 1;
}
int GpuSceneBuilder::add_bsdf(const MskBsdf &b) {
    m_bsdfs.push_back(b);
    return (int) m_bsdfs.size() - 1;
}
int GpuSceneBuilder::spectrum_id(const Texture *t) {
    auto it = m_spectrum_ids.find(t);
    if (it != m_spectrum_ids.end()) return it->second;
    int id = t->describe(*this);
    m_spectrum_ids[t] = id;
    return id;
}
int GpuSceneBuilder::bsdf_id(const BSDF *b) {
    auto it = m_bsdf_ids.find(b);
    if (it != m_bsdf_ids.end()) return it->second;
    int id = b->describe(*this);
    if (id < 0) { // twosided adapter around entry -(id + 2)
        id = -(id + 2);
        m_bsdfs[id].twosided = 1;
    }
    m_bsdf_ids[b] = id;
    return id;
}
int GpuSceneBuilder::medium_id(const Medium *m) {
    if (!m) return -1;
    auto it = m_medium_ids.find(m);
    if (it != m_medium_ids.end()) return it->second;
    MskMedium d{};
    m->describe(*this, d);
    m_media.push_back(d);
    return m_medium_ids[m] = (int) m_media.size() - 1;
}
void GpuSceneBuilder::finish() {
    m_desc.media = m_media.data(); m_desc.nmedia = (uint32_t) m_media.size();
    m_desc.meshes = m_meshes.data(); m_desc.nmeshes = (uint32_t) m_meshes.size();
    m_desc.bsdfs = m_bsdfs.data(); m_desc.nbsdfs = (uint32_t) m_bsdfs.size();
    m_desc.emitters = m_emitters.data(); m_desc.nemitters = (uint32_t) m_emitters.size();
    m_desc.spectra = m_spectra.data(); m_desc.nspectra = (uint32_t) m_spectra.size();
    m_desc.spectrum_tables = m_tables.data(); m_desc.ntable_floats = (uint32_t) m_tables.size();
}

} // namespace misaki
