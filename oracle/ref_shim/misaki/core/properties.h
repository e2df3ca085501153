// Stand-in for include/misaki/core/properties.h: a flat property set filled by the wrapper.  TEST INFRASTRUCTURE.
#pragma once
#include "object.h"
#include <array>
#include <map>
#include <string>
#include <utility>
#include <vector>
namespace misaki {
class Texture;
struct Transform4f;
template <typename T> T msk_ref_make_transform(const float *row_major_4x4); // msk_ref_geometry.h, once transform.h is known
class Properties {
public:
    Properties() {}
    explicit Properties(const std::string &plugin) : plugin_name(plugin) {}
    std::string plugin_name;
    std::map<std::string, std::array<float, 16>> matrices; // row-major 4x4
    template <typename T> T transform(const std::string &n, const T &d) const { // shapes stay in world space: no entry, the default
        auto it = matrices.find(n);
        return it == matrices.end() ? d : msk_ref_make_transform<T>(it->second.data());
    }
    std::map<std::string, float> floats;
    std::map<std::string, bool> bools;
    std::map<std::string, long long> ints;
    std::map<std::string, const void *> pointers;
    std::map<std::string, std::array<float, 3>> colors;
    void set_float(const std::string &n, float v) { floats[n] = v; }
    void set_int(const std::string &n, long long v) { ints[n] = v; }
    void set_pointer(const std::string &n, const void *v) { pointers[n] = v; }
    template <typename C = struct ColorTag> auto color(const std::string &n) const; // defined once Color3 is known (msk_ref_prelude.h)
    std::map<std::string, std::string> strings;
    std::map<std::string, ref<Texture>> textures;
    std::vector<std::pair<std::string, ref<Object>>> children;
    ref<Texture> (*make_default)(float) = nullptr; // how a defaulted texture is built (set by the wrapper)
    std::string id() const { return std::string(); }
    bool has_property(const std::string &n) const { return floats.count(n) || bools.count(n) || strings.count(n) || textures.count(n) || ints.count(n) || pointers.count(n); }
    std::string string(const std::string &n) const { return strings.at(n); }
    std::string string(const std::string &n, const std::string &d) const { auto it = strings.find(n); return it == strings.end() ? d : it->second; }
    float float_(const std::string &n) const { return floats.at(n); }
    float float_(const std::string &n, float d) const { auto it = floats.find(n); return it == floats.end() ? d : it->second; }
    int int_(const std::string &n) const { return (int) ints.at(n); } // properties.h:74 (int)
    int int_(const std::string &n, int d) const { auto it = ints.find(n); return it == ints.end() ? d : (int) it->second; }
    const void *pointer(const std::string &n) const { return pointers.at(n); }
    bool bool_(const std::string &n, bool d) const { auto it = bools.find(n); return it == bools.end() ? d : it->second; }
    ref<Texture> texture(const std::string &n) const { return textures.at(n); }
    ref<Texture> texture(const std::string &n, const ref<Texture> &d) const { auto it = textures.find(n); return it == textures.end() ? d : it->second; }
    ref<Texture> texture(const std::string &n, float d) const { auto it = textures.find(n); return it == textures.end() ? make_default(d) : it->second; }
    const std::vector<std::pair<std::string, ref<Object>>> &objects() const { return children; }
};
} // namespace misaki
