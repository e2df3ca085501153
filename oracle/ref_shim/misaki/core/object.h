// Stand-in for include/misaki/core/object.h (+ class.h): just enough of Object / ref<T> / the class-registration macros
// for the reference's BSDF plugin sources to compile on their own.  TEST INFRASTRUCTURE (oracle/Makefile.ref).
#pragma once
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace misaki {
template <typename T> class ref;
class Object {
public:
    virtual std::vector<ref<Object>> expand() const; // object.h: plugins that stand for other plugins (spectra/d65.cpp)
    virtual ~Object() {}
    virtual std::string id() const { return std::string(); }
    virtual std::string to_string() const { return "Object"; }
};
template <typename T> class ref {
public:
    ref() {}
    ref(T *p) : m_ptr(p) {}
    template <typename T2> ref(const ref<T2> &r) : m_ptr((T2 *) r.get()) {}
    T *get() const { return m_ptr; }
    T *operator->() const { return m_ptr; }
    T &operator*() const { return *m_ptr; }
    operator T *() const { return m_ptr; }
    explicit operator bool() const { return m_ptr != nullptr; }
private:
    T *m_ptr = nullptr; // the stand-in never frees: objects live for the duration of a generator run
};
inline std::vector<ref<Object>> Object::expand() const { return {}; }
template <typename T> std::ostream &operator<<(std::ostream &os, const ref<T> &r) { return os << (r.get() ? r->to_string() : std::string("null")); }
inline std::ostream &operator<<(std::ostream &os, const Object *o) { return os << (o ? o->to_string() : std::string("null")); }
} // namespace misaki
#define MSK_DECLARE_CLASS()
#define MSK_IMPLEMENT_CLASS(...)
#define MSK_REGISTER_INSTANCE(...)
#define MSK_INTERNAL_PLUGIN(...)
