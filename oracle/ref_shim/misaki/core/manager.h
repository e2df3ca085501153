// Stand-in for include/misaki/core/manager.h: a plugin table filled by the wrappers (ref_spectra_wrap.cpp registers
// "regular" and "d65"); a plugin that is not in the table yields a null reference (Shape's default "diffuse" BSDF and
// Scene's default "path" integrator are never used by the pinned code paths).  TEST INFRASTRUCTURE.
#pragma once
#include "object.h"
#include "properties.h"
#include <functional>
#include <map>
#include <string>
namespace misaki {
class InstanceManager {
public:
    static InstanceManager *get() { static InstanceManager m; return &m; }
    std::map<std::string, std::function<Object *(const Properties &)>> table;
    template <typename T> ref<T> create_instance(const Properties &props) {
        auto it = table.find(props.plugin_name);
        return it == table.end() ? ref<T>() : ref<T>(static_cast<T *>(it->second(props)));
    }
};
} // namespace misaki
