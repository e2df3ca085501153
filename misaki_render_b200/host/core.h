// Host-side object system of the B200 backend: the same plugin surface as misaki-render
// (Object / ref<T> / Class / InstanceManager / Properties / xml::load_file), restated without Eigen,
// pugixml, TBB or fmt (none of which exist in this image).  A plugin is a C++ class registered under the
// name the XML `type=` attribute uses, exactly as in the reference:
//
//   MSK_DECLARE_CLASS / MSK_IMPLEMENT_CLASS      reference include/misaki/core/class.h:50-60
//   MSK_REGISTER_INSTANCE(Class, "name")         reference include/misaki/core/manager.h:39-45
//   InstanceManager::create_instance             reference src/librender/manager.cpp:13-37
//   ref<T>, Object                               reference include/misaki/core/object.h:31-148
//   Properties                                   reference include/misaki/core/properties.h, properties.cpp
//   Throw(...) -> std::runtime_error             reference include/misaki/core/logger.h:81-85
#pragma once
#include <array>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

namespace misaki {

// ---------------------------------------------------------------------------------------- logging / errors
enum LogLevel { Trace = 0, Debug, Info, Warn, Error };
void set_log_level(LogLevel l);
void Log(LogLevel level, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
[[noreturn]] void Throw(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
std::string format(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

// ---------------------------------------------------------------------------------------- Object / ref
class Class;
class Properties;
template <typename T> class ref;

class Object {
public:
    Object() = default;
    Object(const Object &) : m_ref_count(0) {}
    void inc_ref() const { ++m_ref_count; }
    void dec_ref(bool dealloc = true) const noexcept {
        if (--m_ref_count == 0 && dealloc) delete this;
    }
    int ref_count() const { return m_ref_count; }
    virtual const Class *clazz() const;
    virtual std::string to_string() const;
    virtual std::vector<ref<Object>> expand() const; // d65 -> regular, reference spectra/d65.cpp:36-50
    static Class *m_class;

protected:
    virtual ~Object();

private:
    mutable std::atomic<int> m_ref_count{ 0 };
};

template <typename T> class ref {
public:
    ref() = default;
    ref(T *p) : m_ptr(p) { if (m_ptr) ((Object *) m_ptr)->inc_ref(); }
    ref(const ref &r) : m_ptr(r.m_ptr) { if (m_ptr) ((Object *) m_ptr)->inc_ref(); }
    ref(ref &&r) noexcept : m_ptr(r.m_ptr) { r.m_ptr = nullptr; }
    ~ref() { if (m_ptr) ((Object *) m_ptr)->dec_ref(); }
    ref &operator=(const ref &r) {
        if (r.m_ptr) ((Object *) r.m_ptr)->inc_ref();
        if (m_ptr) ((Object *) m_ptr)->dec_ref();
        m_ptr = r.m_ptr;
        return *this;
    }
    ref &operator=(ref &&r) noexcept {
        if (&r != this) { if (m_ptr) ((Object *) m_ptr)->dec_ref(); m_ptr = r.m_ptr; r.m_ptr = nullptr; }
        return *this;
    }
    T *operator->() const { return m_ptr; }
    T &operator*() const { return *m_ptr; }
    T *get() const { return m_ptr; }
    operator T *() const { return m_ptr; }
    explicit operator bool() const { return m_ptr != nullptr; }

private:
    T *m_ptr = nullptr;
};

// ---------------------------------------------------------------------------------------- Class registry
class Class {
public:
    using ConstructFunctor = Object *(*) (const Properties &);
    Class(const std::string &name, const std::string &parent, const std::string &alias = "", ConstructFunctor ctor = nullptr);
    const std::string &name() const { return m_name; }
    const std::string &alias() const { return m_alias; }
    const Class *parent() const { return m_parent; }
    bool derives_from(const Class *c) const;
    bool is_constructible() const { return m_construct != nullptr; }
    Object *construct(const Properties &props) const;
    static const Class *for_name(const std::string &name);
    static void static_initialization();

private:
    std::string m_name, m_parent_name, m_alias;
    const Class *m_parent = nullptr;
    ConstructFunctor m_construct;
};

class InstanceManager {
public:
    static InstanceManager *get();
    // plugin `name` (the XML type= string) of kind `class_` (e.g. "BSDF")
    void register_instance(const std::string &name, const Class *class_);
    ref<Object> create_instance(const Properties &props, const Class *class_);
    template <typename T> ref<T> create_instance(const Properties &props) {
        ref<Object> o = create_instance(props, T::m_class);
        return ref<T>(static_cast<T *>(o.get()));
    }
    std::vector<std::string> registered() const;

private:
    std::map<std::string, const Class *> m_classes;
};

#define MSK_CLASS(x) x::m_class
#define MSK_DECLARE_CLASS()                      \
    virtual const Class *clazz() const override; \
public:                                          \
    static Class *m_class;
#define MSK_IMPLEMENT_CLASS(Name, Parent, ...)                    \
    Class *Name::m_class = new Class(#Name, #Parent, ##__VA_ARGS__); \
    const Class *Name::clazz() const { return m_class; }
// constructible plugin: registers the constructor with its Class and the XML type name with the manager
#define MSK_IMPLEMENT_PLUGIN(Name, Parent, Type)                                                               \
    Class *Name::m_class = new Class(#Name, #Parent, "", [](const Properties &p) -> Object * { return new Name(p); }); \
    const Class *Name::clazz() const { return m_class; }                                                       \
    static struct Name##_register_ {                                                                           \
        Name##_register_() { InstanceManager::get()->register_instance(Type, Name::m_class); }                  \
    } Name##_register_instance_;

// ---------------------------------------------------------------------------------------- small math
struct Vector3f { float x = 0, y = 0, z = 0; };
struct Color3 { float r = 0, g = 0, b = 0; };

// reference include/misaki/core/transform.h: a 4x4 matrix together with its inverse
struct Transform4f {
    float m[16], inv[16]; // row-major
    Transform4f();
    explicit Transform4f(const float *matrix); // inverse computed by cofactor expansion (float32)
    Transform4f(const float *matrix, const float *inverse);
    Transform4f operator*(const Transform4f &t) const; // (m * t.m, t.inv * inv), transform.h:100-103
    Vector3f apply_point(const Vector3f &p) const;      // with perspective divide, transform.h:129-136
    Vector3f apply_normal(const Vector3f &n) const;     // inverse transpose, transform.h:138-140
    bool has_nan() const;
    static Transform4f translate(const Vector3f &v);
    static Transform4f scale(const Vector3f &v);
    static Transform4f rotate(const Vector3f &axis, float angle_radians);
    static Transform4f lookat(const Vector3f &origin, const Vector3f &target, const Vector3f &up);
    static Transform4f perspective(float fov, float near_, float far_);
};

// ---------------------------------------------------------------------------------------- Properties
class Texture;
class Properties {
public:
    struct NamedReference { std::string id; };
    enum class Type { Bool, Int, Float, String, Vector3, Color, Transform, NamedReference, Object, Pointer };
    using Value = std::variant<bool, int64_t, float, std::string, Vector3f, Color3, Transform4f, NamedReference, ref<Object>, const void *>;

    Properties() = default;
    explicit Properties(const std::string &instance_name) : m_instance_name(instance_name) {}
    const std::string &instance_name() const { return m_instance_name; }
    void set_instance_name(const std::string &n) { m_instance_name = n; }
    const std::string &id() const { return m_id; }
    void set_id(const std::string &id) { m_id = id; }

    bool has_property(const std::string &name) const { return m_entries.count(name) != 0; }
    Type type(const std::string &name) const;
    void set(const std::string &name, Value v) { m_entries[name] = std::move(v); }
    void set_bool(const std::string &n, bool v) { set(n, v); }
    void set_int(const std::string &n, int64_t v) { set(n, v); }
    void set_float(const std::string &n, float v) { set(n, v); }
    void set_string(const std::string &n, const std::string &v) { set(n, v); }
    void set_vector3(const std::string &n, Vector3f v) { set(n, v); }
    void set_color(const std::string &n, Color3 v) { set(n, v); }
    void set_transform(const std::string &n, const Transform4f &v) { set(n, v); }
    void set_named_reference(const std::string &n, const std::string &id) { set(n, NamedReference{ id }); }
    void set_object(const std::string &n, const ref<Object> &o) { set(n, o); }
    void set_pointer(const std::string &n, const void *p) { set(n, p); }

    // typed getters: throw when missing (no default) or of the wrong type, like properties.cpp:11-54
    bool bool_(const std::string &n) const;
    bool bool_(const std::string &n, bool def) const;
    int64_t int_(const std::string &n) const;
    int64_t int_(const std::string &n, int64_t def) const;
    float float_(const std::string &n) const;
    float float_(const std::string &n, float def) const;
    std::string string(const std::string &n) const;
    std::string string(const std::string &n, const std::string &def) const;
    Vector3f vector3(const std::string &n, Vector3f def) const;
    Color3 color(const std::string &n) const;
    Transform4f transform(const std::string &n, const Transform4f &def) const;
    const void *pointer(const std::string &n) const;

    // std::map order of the property names (properties.cpp:166-176): "_arg_0, _arg_1, _arg_10, _arg_2, ..."
    std::vector<std::pair<std::string, ref<Object>>> objects() const;
    std::vector<std::pair<std::string, std::string>> named_references() const;
    // properties.cpp:194-235
    ref<Texture> texture(const std::string &name) const;
    ref<Texture> texture(const std::string &name, float def_val) const;
    ref<Texture> texture(const std::string &name, const ref<Texture> &def_val) const;

private:
    std::string m_instance_name, m_id;
    std::map<std::string, Value> m_entries;
};

// ---------------------------------------------------------------------------------------- file resolver / xml
class FileResolver {
public:
    void append(const std::string &dir) { m_paths.push_back(dir); }
    void prepend(const std::string &dir) { m_paths.insert(m_paths.begin(), dir); }
    void pop_front() { if (!m_paths.empty()) m_paths.erase(m_paths.begin()); }
    std::string resolve(const std::string &path) const; // first existing <dir>/<path>, else path itself
private:
    std::vector<std::string> m_paths;
};
FileResolver *get_file_resolver();
// The directory of the scene being loaded, searched first FOR THE DURATION OF THAT LOAD (main.cpp:68 prepends it for the
// life of the process, which is one load there; a library that loads many scenes must not let scene B find scene A's files)
struct ScopedSearchPath {
    explicit ScopedSearchPath(const std::string &dir) { get_file_resolver()->prepend(dir); }
    ~ScopedSearchPath() { get_file_resolver()->pop_front(); }
    ScopedSearchPath(const ScopedSearchPath &) = delete;
    ScopedSearchPath &operator=(const ScopedSearchPath &) = delete;
};

namespace xml {
using ParameterList = std::vector<std::pair<std::string, std::string>>;
// reference src/librender/xml.cpp:714-740
ref<Object> load_file(const std::string &filename, ParameterList parameters = {});
ref<Object> load_string(const std::string &text, const std::string &source_id = "<string>", ParameterList parameters = {});
} // namespace xml

namespace string {
std::vector<std::string> tokenize(const std::string &s, const std::string &delim = ", ", bool include_empty = false);
std::string to_lower(std::string s);
bool starts_with(const std::string &s, const std::string &prefix);
} // namespace string

} // namespace misaki
