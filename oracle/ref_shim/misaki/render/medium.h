// Stand-in for include/misaki/render/medium.h: the pinned code paths run without participating media.  TEST INFRASTRUCTURE.
#pragma once
#include "msk_ref_prelude.h"
#include <misaki/core/object.h>
#include <misaki/core/ray.h>
namespace misaki {
class Medium : public Object {
public:
    virtual Spectrum eval_transmittance(const Ray &) const { throw 1; }
};
} // namespace misaki
