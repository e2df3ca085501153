// C entry point over the reference's OWN colour-to-spectrum plugins: src/librender/srgb.cpp (srgb_model_fetch on
// ext/rgb2spec), spectra/srgb.cpp, spectra/srgb_d65.cpp and spectra/d65.cpp (which expands into spectra/regular.cpp through
// the plugin manager), #included from where they lie over the stand-ins under oracle/ref_shim/.  These are what an <rgb>
// tag becomes (xml.cpp:269-277).  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#include <misaki/core/manager.h>
#include <misaki/render/interaction.h>
#include <misaki/render/texture.h>
#include <srgb.cpp>
#include <spectra/srgb.cpp>
#include <spectra/srgb_d65.cpp>
#include <spectra/d65.cpp>

using namespace misaki;
misaki::Object *msk_ref_make_regular(const misaki::Properties &p); // spectra/regular.cpp lives in ref_plugins_wrap.cpp

static void register_expansions() {
    InstanceManager *mgr = InstanceManager::get();
    mgr->table["regular"] = [](const Properties &p) -> Object * { return msk_ref_make_regular(p); };
    mgr->table["d65"] = [](const Properties &p) -> Object * { return new D65Spectrum(p); };
}
// What create_texture_from_rgb (xml.cpp:269-277) builds for an <rgb> tag: "srgb", or "srgb_d65" inside an <emitter>, with the
// colour as the only property.  Used by ref_path_wrap.cpp to put the reference's own colour textures into its scenes.
misaki::Texture *msk_ref_make_rgb_texture(const float rgb[3], bool within_emitter) {
    register_expansions();
    Properties p(within_emitter ? "srgb_d65" : "srgb");
    p.colors["color"] = { rgb[0], rgb[1], rgb[2] };
    if (within_emitter) return new SRGBEmitterSpectrum(p);
    return new SRGBReflectanceSpectrum(p);
}

// kind: 0 "srgb" (reflectance <rgb>), 1 "srgb_d65" (<rgb> inside an emitter, with "scale"), 2 "d65" (scale) expanded
extern "C" int ref_colour_spectrum(int kind, const float rgb[3], float scale, const float wl[4], float out[4]) {
    try {
        register_expansions();
        Properties p;
        p.colors["color"] = { rgb[0], rgb[1], rgb[2] };
        p.floats["scale"] = scale;
        ref<Texture> t;
        if (kind == 0) t = new SRGBReflectanceSpectrum(p);
        else if (kind == 1) t = new SRGBEmitterSpectrum(p);
        else { D65Spectrum d(p); t = (Texture *) d.expand().at(0).get(); }
        SceneInteraction si;
        si.wavelengths = Wavelength(wl[0], wl[1], wl[2], wl[3]);
        Spectrum v = t->eval(si);
        for (int i = 0; i < 4; ++i) out[i] = v.coeff(i);
        return 0;
    } catch (...) { return -2; }
}
