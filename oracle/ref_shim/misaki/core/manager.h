// Stand-in for include/misaki/core/manager.h: no plugin registry; a defaulted plugin instance is a null reference
// (Shape's default "diffuse" BSDF is never evaluated by the pinned code paths).  TEST INFRASTRUCTURE.
#pragma once
#include "object.h"
namespace misaki {
class Properties;
class InstanceManager {
public:
    static InstanceManager *get() { static InstanceManager m; return &m; }
    template <typename T> ref<T> create_instance(const Properties &) { return ref<T>(); }
};
} // namespace misaki
