// Stand-in for include/misaki/render/sensor.h (+ film.h, imageblock.h, image.h): Scene only stores its sensor; camera rays
// are supplied by the caller of the pinned PathTracer::sample.  TEST INFRASTRUCTURE.
#pragma once
#include <misaki/core/object.h>
#include <misaki/render/sampler.h>
namespace misaki {
class Film : public Object {};
class ImageBlock : public Object {};
class Sensor : public Object { public: const Medium *medium() const { return nullptr; } };
} // namespace misaki
