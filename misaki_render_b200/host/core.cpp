// Object system, Properties, Transform4f, logging: see core.h for the reference files each part mirrors.
#include "core.h"
#include "render.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sys/stat.h>

namespace misaki {

// ---------------------------------------------------------------------------------------- logging
static LogLevel g_level = Info;
void set_log_level(LogLevel l) { g_level = l; }

static std::string vformat(const char *fmt, va_list ap) {
    va_list ap2;
    va_copy(ap2, ap);
    int n = vsnprintf(nullptr, 0, fmt, ap2);
    va_end(ap2);
    std::string s((size_t) std::max(n, 0), '\0');
    vsnprintf(s.data(), s.size() + 1, fmt, ap);
    return s;
}
std::string format(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::string s = vformat(fmt, ap);
    va_end(ap);
    return s;
}
void Log(LogLevel level, const char *fmt, ...) {
    if (level < g_level) return;
    static const char *names[] = { "trace", "debug", "info", "warn", "error" };
    va_list ap;
    va_start(ap, fmt);
    std::string s = vformat(fmt, ap);
    va_end(ap);
    fprintf(stderr, "[%s] %s\n", names[level], s.c_str());
}
void Throw(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::string s = vformat(fmt, ap);
    va_end(ap);
    throw std::runtime_error(s);
}

// ---------------------------------------------------------------------------------------- strings
namespace string {
std::vector<std::string> tokenize(const std::string &s, const std::string &delim, bool include_empty) {
    std::vector<std::string> out;
    size_t last = 0, pos;
    while (true) {
        pos = s.find_first_of(delim, last);
        size_t end = pos == std::string::npos ? s.size() : pos;
        if (end != last || include_empty) out.push_back(s.substr(last, end - last));
        if (pos == std::string::npos) break;
        last = pos + 1;
    }
    return out;
}
std::string to_lower(std::string s) {
    for (auto &c : s) c = (char) std::tolower((unsigned char) c);
    return s;
}
bool starts_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
} // namespace string

// ---------------------------------------------------------------------------------------- Object / Class
static std::map<std::string, Class *> &class_map() {
    static std::map<std::string, Class *> m;
    return m;
}
namespace xml { void register_class(const Class *c); }

Class::Class(const std::string &name, const std::string &parent, const std::string &alias, ConstructFunctor ctor)
    : m_name(name), m_parent_name(parent), m_alias(alias.empty() ? name : alias), m_construct(ctor) {
    class_map()[name] = this;
    if (!alias.empty()) xml::register_class(this); // an alias names an XML object tag ("bsdf", "shape", ...)
}
bool Class::derives_from(const Class *c) const {
    for (const Class *k = this; k; k = k->m_parent)
        if (k == c) return true;
    return false;
}
Object *Class::construct(const Properties &props) const {
    if (!m_construct) Throw("RTTI error: Attempted to construct a non-constructible class \"%s\"!", m_name.c_str());
    return m_construct(props);
}
const Class *Class::for_name(const std::string &name) {
    auto it = class_map().find(name);
    return it == class_map().end() ? nullptr : it->second;
}
void Class::static_initialization() { // class.cpp:80-84: link parents by name
    for (auto &kv : class_map()) {
        Class *c = kv.second;
        if (c->m_parent || c->m_parent_name.empty()) continue;
        auto it = class_map().find(c->m_parent_name);
        if (it != class_map().end()) c->m_parent = it->second;
    }
}

Class *Object::m_class = new Class("Object", "");
const Class *Object::clazz() const { return m_class; }
std::string Object::to_string() const { return format("%s[%p]", clazz()->name().c_str(), (const void *) this); }
std::vector<ref<Object>> Object::expand() const { return {}; }
Object::~Object() {}

InstanceManager *InstanceManager::get() {
    static InstanceManager im;
    return &im;
}
void InstanceManager::register_instance(const std::string &name, const Class *class_) { m_classes[name] = class_; }
std::vector<std::string> InstanceManager::registered() const {
    std::vector<std::string> r;
    for (auto &kv : m_classes) r.push_back(kv.first);
    return r;
}
ref<Object> InstanceManager::create_instance(const Properties &props, const Class *class_) {
    Class::static_initialization();
    auto it = m_classes.find(props.instance_name()); // manager.cpp:18-20
    if (it == m_classes.end()) Throw("Plugin \"%s\" not found!", props.instance_name().c_str());
    const Class *c = it->second;
    if (class_ && !c->derives_from(class_))
        Throw("Type mismatch when loading plugin \"%s\": Expected an instance of type \"%s\", got an instance of type \"%s\"",
              props.instance_name().c_str(), class_->name().c_str(), c->name().c_str());
    return ref<Object>(c->construct(props));
}

// ---------------------------------------------------------------------------------------- Transform4f
static void mat_mul(const float *a, const float *b, float *out) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    memcpy(out, r, sizeof(r));
}
static void mat_inverse(const float *m, float *out) { // cofactor expansion, float32
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float inv_det = 1.f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * inv_det;
}

Transform4f::Transform4f() {
    for (int i = 0; i < 16; ++i) m[i] = inv[i] = (i % 5 == 0) ? 1.f : 0.f;
}
Transform4f::Transform4f(const float *matrix) {
    memcpy(m, matrix, sizeof(m));
    mat_inverse(m, inv);
}
Transform4f::Transform4f(const float *matrix, const float *inverse) {
    memcpy(m, matrix, sizeof(m));
    memcpy(inv, inverse, sizeof(inv));
}
Transform4f Transform4f::operator*(const Transform4f &t) const {
    Transform4f r;
    mat_mul(m, t.m, r.m);
    mat_mul(t.inv, inv, r.inv);
    return r;
}
Vector3f Transform4f::apply_point(const Vector3f &p) const {
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = m[i * 4] * p.x + m[i * 4 + 1] * p.y + m[i * 4 + 2] * p.z + m[i * 4 + 3] * 1.f;
    return Vector3f{ r[0] / r[3], r[1] / r[3], r[2] / r[3] };
}
Vector3f Transform4f::apply_normal(const Vector3f &n) const { // (inv^T)_3x3 * n
    return Vector3f{ inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
                     inv[2] * n.x + inv[6] * n.y + inv[10] * n.z };
}
bool Transform4f::has_nan() const {
    for (float v : m) if (std::isnan(v)) return true;
    return false;
}
Transform4f Transform4f::translate(const Vector3f &v) {
    float a[16] = { 1, 0, 0, v.x, 0, 1, 0, v.y, 0, 0, 1, v.z, 0, 0, 0, 1 };
    return Transform4f(a);
}
Transform4f Transform4f::scale(const Vector3f &v) {
    float a[16] = { v.x, 0, 0, 0, 0, v.y, 0, 0, 0, 0, v.z, 0, 0, 0, 0, 1 };
    return Transform4f(a);
}
static Vector3f normalized(Vector3f v) {
    float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return n > 0.f ? Vector3f{ v.x / n, v.y / n, v.z / n } : v;
}
Transform4f Transform4f::rotate(const Vector3f &axis_, float angle) { // transform.h:163-167 (Eigen::AngleAxisf: Rodrigues, radians)
    Vector3f a = normalized(axis_);
    float s = std::sin(angle), c = std::cos(angle), t = 1.f - c;
    float r[16] = { t * a.x * a.x + c,       t * a.x * a.y - s * a.z, t * a.x * a.z + s * a.y, 0,
                    t * a.x * a.y + s * a.z, t * a.y * a.y + c,       t * a.y * a.z - s * a.x, 0,
                    t * a.x * a.z - s * a.y, t * a.y * a.z + s * a.x, t * a.z * a.z + c,       0,
                    0, 0, 0, 1 };
    float ri[16]; // the inverse of a rotation is its transpose
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) ri[i * 4 + j] = r[j * 4 + i];
    return Transform4f(r, ri);
}
static Vector3f cross(Vector3f a, Vector3f b) { return Vector3f{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
Transform4f Transform4f::lookat(const Vector3f &origin, const Vector3f &target, const Vector3f &up) { // transform.h:170-179
    Vector3f dir = normalized(Vector3f{ target.x - origin.x, target.y - origin.y, target.z - origin.z });
    Vector3f left = normalized(cross(normalized(up), dir));
    Vector3f new_up = normalized(cross(dir, left));
    float a[16] = { left.x, new_up.x, dir.x, origin.x, left.y, new_up.y, dir.y, origin.y, left.z, new_up.z, dir.z, origin.z, 0, 0, 0, 1 };
    return Transform4f(a);
}
Transform4f Transform4f::perspective(float fov, float near_, float far_) { // transform.h:181-188
    float recip = 1.0f / (far_ - near_);
    float cot = 1.0f / std::tan(fov / 2.0f * (float) (M_PI / 180.0));
    float a[16] = { cot, 0, 0, 0, 0, cot, 0, 0, 0, 0, far_ * recip, -near_ * far_ * recip, 0, 0, 1, 0 };
    return Transform4f(a);
}

// ---------------------------------------------------------------------------------------- Properties
Properties::Type Properties::type(const std::string &name) const {
    auto it = m_entries.find(name);
    if (it == m_entries.end()) Throw("type(): Could not find property named \"%s\"!", name.c_str());
    return (Type) it->second.index();
}

#define MSK_GETTER(FuncName, CType, VariantType, TypeName)                                                       \
    CType Properties::FuncName(const std::string &name) const {                                                  \
        auto it = m_entries.find(name);                                                                          \
        if (it == m_entries.end()) Throw("Property \"%s\" has not been specified!", name.c_str());               \
        if (!std::holds_alternative<VariantType>(it->second))                                                    \
            Throw("The property \"%s\" has the wrong type (expected <" TypeName ">).", name.c_str());             \
        return (CType) std::get<VariantType>(it->second);                                                        \
    }
#define MSK_GETTER_DEF(FuncName, CType, VariantType, TypeName)                                                   \
    CType Properties::FuncName(const std::string &name, CType def) const {                                       \
        auto it = m_entries.find(name);                                                                          \
        if (it == m_entries.end()) return def;                                                                   \
        if (!std::holds_alternative<VariantType>(it->second))                                                    \
            Throw("The property \"%s\" has the wrong type (expected <" TypeName ">).", name.c_str());             \
        return (CType) std::get<VariantType>(it->second);                                                        \
    }
MSK_GETTER(bool_, bool, bool, "boolean")
MSK_GETTER_DEF(bool_, bool, bool, "boolean")
MSK_GETTER(int_, int64_t, int64_t, "integer")
MSK_GETTER_DEF(int_, int64_t, int64_t, "integer")
MSK_GETTER(float_, float, float, "float")
MSK_GETTER_DEF(float_, float, float, "float")
MSK_GETTER(color, Color3, Color3, "rgb")
MSK_GETTER(pointer, const void *, const void *, "pointer")
std::string Properties::string(const std::string &name) const {
    auto it = m_entries.find(name);
    if (it == m_entries.end()) Throw("Property \"%s\" has not been specified!", name.c_str());
    if (!std::holds_alternative<std::string>(it->second)) Throw("The property \"%s\" has the wrong type (expected <string>).", name.c_str());
    return std::get<std::string>(it->second);
}
std::string Properties::string(const std::string &name, const std::string &def) const {
    return has_property(name) ? string(name) : def;
}
Vector3f Properties::vector3(const std::string &name, Vector3f def) const {
    auto it = m_entries.find(name);
    if (it == m_entries.end()) return def;
    if (!std::holds_alternative<Vector3f>(it->second)) Throw("The property \"%s\" has the wrong type (expected <vector>).", name.c_str());
    return std::get<Vector3f>(it->second);
}
Transform4f Properties::transform(const std::string &name, const Transform4f &def) const {
    auto it = m_entries.find(name);
    if (it == m_entries.end()) return def;
    if (!std::holds_alternative<Transform4f>(it->second)) Throw("The property \"%s\" has the wrong type (expected <transform>).", name.c_str());
    return std::get<Transform4f>(it->second);
}
std::vector<std::pair<std::string, ref<Object>>> Properties::objects() const {
    std::vector<std::pair<std::string, ref<Object>>> r;
    for (auto &kv : m_entries)
        if (std::holds_alternative<ref<Object>>(kv.second)) r.emplace_back(kv.first, std::get<ref<Object>>(kv.second));
    return r;
}
std::vector<std::pair<std::string, std::string>> Properties::named_references() const {
    std::vector<std::pair<std::string, std::string>> r;
    for (auto &kv : m_entries)
        if (std::holds_alternative<NamedReference>(kv.second)) r.emplace_back(kv.first, std::get<NamedReference>(kv.second).id);
    return r;
}

ref<Texture> Properties::texture(const std::string &name) const { // properties.cpp:194-218
    if (!has_property(name)) Throw("Property %s has not been specified!", name.c_str());
    Type t = type(name);
    if (t == Type::Object) {
        ref<Object> o = std::get<ref<Object>>(m_entries.at(name));
        if (!o->clazz()->derives_from(MSK_CLASS(Texture)))
            Throw("The property \"%s\" has the wrong type (expected  <spectrum> or <texture>).", name.c_str());
        return ref<Texture>(static_cast<Texture *>(o.get()));
    } else if (t == Type::Float) {
        Properties p("uniform");
        p.set_float("value", float_(name));
        return InstanceManager::get()->create_instance<Texture>(p);
    }
    Throw("The property \"%s\" has the wrong type (expected  <spectrum> or <texture>).", name.c_str());
}
ref<Texture> Properties::texture(const std::string &name, const ref<Texture> &def_val) const {
    return has_property(name) ? texture(name) : def_val;
}
ref<Texture> Properties::texture(const std::string &name, float def_val) const {
    if (!has_property(name)) {
        // properties.cpp:226-233 builds an "srgb" texture with key "value", which SRGBReflectanceSpectrum (key
        // "color", spectra/srgb.cpp:16) rejects, so every defaulted texture throws in the reference.  The evident
        // intent -- a gray srgb texture of the default value -- is implemented here.
        Properties p("srgb");
        p.set_color("color", Color3{ def_val, def_val, def_val });
        return InstanceManager::get()->create_instance<Texture>(p);
    }
    return texture(name);
}

// ---------------------------------------------------------------------------------------- FileResolver
static bool file_exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
std::string FileResolver::resolve(const std::string &path) const {
    if (!path.empty() && path[0] == '/') return path;
    for (auto &d : m_paths) {
        std::string c = d.empty() ? path : d + "/" + path;
        if (file_exists(c)) return c;
    }
    return path;
}
FileResolver *get_file_resolver() {
    static FileResolver fr;
    return &fr;
}

} // namespace misaki
