// Stand-in for include/misaki/render/emitter.h: Shape only stores and notifies its emitter.  TEST INFRASTRUCTURE.
#pragma once
#include <misaki/core/object.h>
namespace misaki {
class Shape;
class Emitter : public Object { public: void set_shape(Shape *) {} };
} // namespace misaki
