// C entry point over the reference's OWN BSDF plugin source src/librender/bsdfs/diffuse.cpp -- the one BSDF its build
// compiles (src/librender/CMakeLists.txt:61-67) -- #included from where it lies and compiled against the Eigen stand-in
// and the object-system stand-ins under oracle/ref_shim/ (oracle/Makefile.ref -> oracle/_ref/libmisaki_ref_math.so).
// roughconductor / roughdielectric / dielectric / twosided.cpp are commented out of that build for a reason: they mix
// the 3-channel Color3 of the previous RGB pipeline with the 4-wavelength Spectrum (`Color3 F = 1.f`, `Spectrum = Color3`,
// `Spectrum::setConstant`) and do not compile against the current headers with any Eigen; tried here, same errors.
// TEST INFRASTRUCTURE: tools/gen_golden_ref_math.py calls ref_bsdf to produce golden sample / eval / pdf vectors.
//
// Not taken from the reference: the textures (a constant spectrum stands for every spectral parameter, so the BSDF
// arithmetic is what is compared, not the spectra), and the trivial Texture base-class members that live in texture.cpp
// next to code needing the plugin manager.
#include "msk_ref_prelude.h"
#include <misaki/render/bsdf.h>
#include <misaki/render/texture.h>
#include "ref_wrap_common.h"

namespace misaki {
Texture::Texture(const Properties &props) : m_id(props.id()) {}
Texture::~Texture() {}
float Texture::eval_1(const SceneInteraction &) const { throw 1; }
Spectrum Texture::eval(const SceneInteraction &) const { throw 1; }
Color3 Texture::eval_3(const SceneInteraction &) const { throw 1; }
float Texture::mean() const { throw 1; }

} // namespace misaki

#include <bsdf.cpp>          // BSDF::BSDF, ~BSDF, id(), SceneInteraction::bsdf(ray)
#include <bsdfs/diffuse.cpp>

using namespace misaki;

// type: 0 diffuse (MskBsdfType).  params[0]: reflectance.
// out_sample = [wo.xyz, pdf, eta, sampled_type, weight0..3] of sample(ctx, si, smp[0], smp[1..2]); out_eval / out_pdf of wo.
extern "C" int ref_bsdf(int type, const float params[10], const float wi[3], const float wl[4], const float smp[3], const float wo_in[3],
                        float out_sample[10], float out_eval[4], float *out_pdf) {
    try {
        Properties p;
        p.make_default = make_const;
        BSDF *b = nullptr;
        if (type == 0) {
            p.textures["reflectance"] = make_const(params[0]);
            b = new SmoothDiffuse(p);
        } else
            return -1;
        SceneInteraction si;
        si.t = 1.f;
        si.wi = Eigen::Vector3f(wi[0], wi[1], wi[2]);
        si.wavelengths = Wavelength(wl[0], wl[1], wl[2], wl[3]);
        si.uv = Eigen::Vector2f(0.f, 0.f);
        BSDFContext ctx;
        auto [bs, w] = b->sample(ctx, si, smp[0], Eigen::Vector2f(smp[1], smp[2]));
        out_sample[0] = bs.wo.x(); out_sample[1] = bs.wo.y(); out_sample[2] = bs.wo.z(); out_sample[3] = bs.pdf; out_sample[4] = bs.eta;
        out_sample[5] = (float) bs.sampled_type;
        for (int i = 0; i < 4; ++i) out_sample[6 + i] = w.coeff(i);
        Eigen::Vector3f wo(wo_in[0], wo_in[1], wo_in[2]);
        Spectrum e = b->eval(ctx, si, wo);
        for (int i = 0; i < 4; ++i) out_eval[i] = e.coeff(i);
        *out_pdf = b->pdf(ctx, si, wo);
        return 0;
    } catch (...) {
        return -2;
    }
}
