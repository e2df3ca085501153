// C entry points over the reference's OWN film: include/misaki/render/film.h, src/librender/film.cpp and
// src/librender/films/hdrfilm.cpp, #included from where they lie.  HDRFilm::prepare / put serve the render loop of
// ref_render_wrap.cpp; HDRFilm::image (hdrfilm.cpp:48-90: XYZ -> linear sRGB, division by the W channel, alpha = A / W,
// AOV channels / W) is exported on its own so that the develop step of the oracle AND of the product's host front-end can be
// compared with it.  Image (OpenImageIO in the reference) is a plain buffer stand-in: HDRFilm::develop's file output is not
// part of the pinned build.  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#include <iostream>
#include <mutex>
#include <misaki/core/image.h>
#include <misaki/core/manager.h>
#include <misaki/core/properties.h>
#include <misaki/core/spectrum.h>
#include <misaki/core/string.h>
#include <misaki/render/film.h>
#include <misaki/render/imageblock.h>
#include <film.cpp>
// HDRFilm is a final class that keeps its storage block protected and offers no accessor besides image(); the comparison
// of whole XYZAW films needs the raw block, so the member access of this ONE class is opened for the wrapper.
#define protected public
#include <films/hdrfilm.cpp>
#undef protected

using namespace misaki;

Film *msk_ref_make_hdrfilm(int W, int H, const ReconstructionFilter *filter) {
    Properties props("hdrfilm");
    props.ints["width"] = W;
    props.ints["height"] = H;
    if (filter) props.children.push_back({ "rfilter", ref<Object>((Object *) filter) }); // film.cpp:23-33
    return new HDRFilm(props);
}
const float *msk_ref_hdrfilm_storage(Film *film) { return static_cast<HDRFilm *>(film)->m_storage->data().data(); }

// storage: H x W x nch accumulated film (X, Y, Z, A, W, aovs...); out: H x W x (nch - 1) developed image (R, G, B, A, aovs...)
extern "C" int ref_hdrfilm_image(const float *storage, int W, int H, int nch, float *out) {
    try {
        Properties fp;
        fp.floats["stddev"] = 0.5f;
        Film *film = msk_ref_make_hdrfilm(W, H, nullptr);
        std::vector<std::string> channels = { "X", "Y", "Z", "A", "W" };
        for (int c = 5; c < nch; ++c) channels.push_back("aov" + std::to_string(c));
        film->prepare(channels);
        ImageBlock *block = new ImageBlock(Eigen::Vector2i(W, H), (size_t) nch); // no filter: no border
        block->set_offset(Eigen::Vector2i(0, 0));
        memcpy(block->data().data(), storage, sizeof(float) * (size_t) W * H * nch);
        film->put(block); // cleared storage + block
        std::shared_ptr<Image> img = film->image();
        if ((int) img->channel_names().size() != nch - 1) return -3;
        memcpy(out, img->pixels().data(), sizeof(float) * img->pixels().size());
        return 0;
    } catch (...) { return -2; }
}
