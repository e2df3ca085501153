// C entry point over the reference's OWN src/librender/imageblock.cpp (ImageBlock::put(pos, value): the filtered splat;
// ImageBlock::put(block) / accumulate_2d: the merge into the film; BlockGenerator: the spiral) driven the way
// SamplingIntegrator::render drives them (integrator.cpp:44-75, films/hdrfilm.cpp:28-46), #included from where it lies
// over the stand-ins under oracle/ref_shim/ (oracle/Makefile.ref).  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#include <misaki/render/rfilter.h>
#include <misaki/render/imageblock.h>
#include <filters/gaussian.cpp> // class GaussianFilter (in-class members only; also in ref_plugins_wrap.cpp)
#include <imageblock.cpp>

using namespace misaki;

// samples: n x (2 + nch) floats = position (pixel units, integrator.cpp:108), channel values.  Blocks are visited in the
// generator's spiral order; within a block the samples whose pixel lies in it are put in array order.  film_out: H x W x nch.
extern "C" int ref_film_accumulate(float stddev, int W, int H, int nch, int block_size, const float *samples, size_t n, float *film_out,
                                   int *block_order /* nblocks x 4: ox oy sx sy, may be null */) {
    try {
        Properties fp;
        fp.floats["stddev"] = stddev;
        GaussianFilter *filter = new GaussianFilter(fp);
        ImageBlock *storage = new ImageBlock(Eigen::Vector2i(W, H), (size_t) nch); // HDRFilm::prepare, hdrfilm.cpp:35-37
        storage->set_offset(Eigen::Vector2i(0, 0));
        storage->clear();
        BlockGenerator *gen = new BlockGenerator(Eigen::Vector2i(W, H), Eigen::Vector2i::Zero(), block_size);
        ImageBlock *block = new ImageBlock(Eigen::Vector2i::Constant(block_size), (size_t) nch, filter, false);
        const size_t stride = 2 + (size_t) nch;
        for (size_t b = 0; b < gen->block_count(); ++b) {
            auto [offset, size, id] = gen->next_block();
            if (block_order) { block_order[4 * b] = offset.x(); block_order[4 * b + 1] = offset.y(); block_order[4 * b + 2] = size.x(); block_order[4 * b + 3] = size.y(); }
            block->set_offset(offset);
            block->set_size(size);
            block->clear(); // render_block, integrator.cpp:84
            for (size_t i = 0; i < n; ++i) {
                const float *s = samples + i * stride;
                int px = (int) std::floor(s[0]), py = (int) std::floor(s[1]);
                if (px < offset.x() || py < offset.y() || px >= offset.x() + size.x() || py >= offset.y() + size.y()) continue;
                block->put(Eigen::Vector2f(s[0], s[1]), s + 2);
            }
            storage->put(block);
        }
        memcpy(film_out, storage->data().data(), sizeof(float) * (size_t) W * H * nch);
        return 0;
    } catch (...) { return -2; }
}
