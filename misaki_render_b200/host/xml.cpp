// Scene-description loader: accepts misaki's XML scene files unchanged.
// Follows reference src/librender/xml.cpp (parse_xml :344-674, instantiate_node :676-710, load_file :714-740)
// on top of a small self-contained XML reader (pugixml is not available in this image).
//
// Behaviours of the reference that are kept on purpose:
//   * children without a name attribute are called _arg_<n> and objects() later returns them in std::map
//     order (_arg_0, _arg_1, _arg_10, _arg_2, ...), which fixes Scene::m_shapes order == geomID
//   * <rgb> becomes an "srgb" texture, or "srgb_d65" inside an <emitter> (:269-277,555-559)
// Front-end completion: <boolean>, <rotate>, <include>, <alias>, <default> are registered tags
// with no case in the reference's switch (:74-90 vs :421-662), i.e. silently dropped there.  They are implemented
// here with the semantics of the Mitsuba 2 loader the reference's xml.cpp descends from:
//   <boolean name value="true|false">            -> Properties::set_bool
//   <rotate x y z angle> | <rotate value angle>  -> Transform4f::rotate(axis, angle in DEGREES) (transform.h:163-167)
//   <default name value>                         -> adds a $name parameter unless the caller already passed one
//   <include filename>                           -> parses that file (relative to the including file); the children of
//                                                   its <scene> root (or its single root object) become children here
//   <alias id as>                                -> a second id for an already declared object
#include "core.h"
#include "render.h"

#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <unordered_map>

namespace misaki {
namespace xml {

// ------------------------------------------------------------------------------------------- mini DOM
namespace {

struct Node {
    enum Kind { Element, Comment, Declaration } kind = Element;
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<Node> children;
    int line = 0;
    const std::string *attr(const char *n) const {
        for (auto &a : attrs) if (a.first == n) return &a.second;
        return nullptr;
    }
    std::string value(const char *n) const { auto *a = attr(n); return a ? *a : std::string(); }
    void append(const std::string &n, const std::string &v) { attrs.emplace_back(n, v); }
    void remove(const char *n) {
        for (size_t i = 0; i < attrs.size(); ++i) if (attrs[i].first == n) { attrs.erase(attrs.begin() + i); return; }
    }
};

class Reader {
public:
    Reader(const std::string &text, const std::string &id) : s(text), src(id) {}
    Node parse_document() {
        Node root;
        bool have_root = false;
        for (;;) {
            skip_ws();
            if (pos >= s.size()) break;
            if (s[pos] != '<') fail("unexpected content");
            Node n = parse_node();
            if (n.kind == Node::Element) {
                if (have_root) fail("multiple root elements");
                root = std::move(n);
                have_root = true;
            }
        }
        if (!have_root) fail("no root element");
        return root;
    }

private:
    const std::string &s;
    std::string src;
    size_t pos = 0;
    int line = 1;

    [[noreturn]] void fail(const char *what) { Throw("Error while loading \"%s\" (at line %d): %s", src.c_str(), line, what); }
    void adv(size_t n = 1) {
        for (size_t i = 0; i < n && pos < s.size(); ++i) { if (s[pos] == '\n') ++line; ++pos; }
    }
    void skip_ws() { while (pos < s.size() && isspace((unsigned char) s[pos])) adv(); }
    bool starts(const char *t) const { return s.compare(pos, strlen(t), t) == 0; }
    static bool name_char(char c) { return isalnum((unsigned char) c) || c == '_' || c == '-' || c == ':' || c == '.'; }
    std::string parse_name() {
        size_t b = pos;
        while (pos < s.size() && name_char(s[pos])) adv();
        if (b == pos) fail("expected a name");
        return s.substr(b, pos - b);
    }
    static std::string decode(const std::string &v) {
        std::string o;
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i] != '&') { o += v[i]; continue; }
            static const std::pair<const char *, char> ents[] = { { "&lt;", '<' }, { "&gt;", '>' }, { "&amp;", '&' }, { "&quot;", '"' }, { "&apos;", '\'' } };
            bool hit = false;
            for (auto &e : ents)
                if (v.compare(i, strlen(e.first), e.first) == 0) { o += e.second; i += strlen(e.first) - 1; hit = true; break; }
            if (!hit) o += v[i];
        }
        return o;
    }
    Node parse_node() {
        Node n;
        n.line = line;
        if (starts("<!--")) {
            size_t e = s.find("-->", pos + 4);
            if (e == std::string::npos) fail("unterminated comment");
            n.kind = Node::Comment;
            adv(e + 3 - pos);
            return n;
        }
        if (starts("<?")) {
            size_t e = s.find("?>", pos + 2);
            if (e == std::string::npos) fail("unterminated declaration");
            n.kind = Node::Declaration;
            adv(e + 2 - pos);
            return n;
        }
        if (starts("<!")) { // DOCTYPE etc.
            size_t e = s.find('>', pos);
            if (e == std::string::npos) fail("unterminated markup declaration");
            n.kind = Node::Declaration;
            adv(e + 1 - pos);
            return n;
        }
        adv(); // '<'
        n.name = parse_name();
        for (;;) {
            skip_ws();
            if (pos >= s.size()) fail("unexpected end of file inside a tag");
            if (starts("/>")) { adv(2); return n; }
            if (s[pos] == '>') { adv(); break; }
            std::string an = parse_name();
            skip_ws();
            if (pos >= s.size() || s[pos] != '=') fail("expected '=' after attribute name");
            adv();
            skip_ws();
            if (pos >= s.size() || (s[pos] != '"' && s[pos] != '\'')) fail("expected a quoted attribute value");
            char q = s[pos];
            adv();
            size_t b = pos;
            while (pos < s.size() && s[pos] != q) adv();
            if (pos >= s.size()) fail("unterminated attribute value");
            n.attrs.emplace_back(an, decode(s.substr(b, pos - b)));
            adv();
        }
        // children until </name>
        for (;;) {
            skip_ws();
            if (pos >= s.size()) fail("unexpected end of file (missing closing tag)");
            if (starts("</")) {
                adv(2);
                std::string cn = parse_name();
                if (cn != n.name) fail("mismatched closing tag");
                skip_ws();
                if (pos >= s.size() || s[pos] != '>') fail("malformed closing tag");
                adv();
                return n;
            }
            if (s[pos] != '<') fail("unexpected content"); // text nodes: the reference rejects them too (:364-365)
            n.children.push_back(parse_node());
        }
    }
};

// ------------------------------------------------------------------------------------------- tags
enum class Tag { Boolean, Integer, Float, String, Vector, Transform, Translate, Matrix, Rotate, Scale, LookAt, Object,
                 NamedReference, RGB, Spectrum, Include, Alias, Default, Invalid };

std::unordered_map<std::string, Tag> &tags() {
    static std::unordered_map<std::string, Tag> t = {
        { "boolean", Tag::Boolean }, { "integer", Tag::Integer }, { "float", Tag::Float }, { "string", Tag::String },
        { "vector", Tag::Vector }, { "transform", Tag::Transform }, { "translate", Tag::Translate }, { "matrix", Tag::Matrix },
        { "rotate", Tag::Rotate }, { "scale", Tag::Scale }, { "lookat", Tag::LookAt }, { "ref", Tag::NamedReference },
        { "rgb", Tag::RGB }, { "spectrum", Tag::Spectrum }, { "include", Tag::Include }, { "alias", Tag::Alias },
        { "default", Tag::Default },
    };
    return t;
}
std::unordered_map<std::string, const Class *> &tag_class() {
    static std::unordered_map<std::string, const Class *> t;
    return t;
}

struct Instance {
    Properties props;
    const Class *clazz = nullptr;
    int line = 0;
    ref<Object> object;
};
struct Context {
    std::string src;
    std::unordered_map<std::string, Instance> instances;
    std::unordered_map<std::string, std::string> aliases; // <alias as="..."> -> id
    Transform4f transform;
    size_t id_counter = 0;
    int include_depth = 0;
};

[[noreturn]] void node_error(const Context &ctx, const Node &n, const std::string &msg) {
    Throw("Error while loading \"%s\" (near line %d): %s", ctx.src.c_str(), n.line, msg.c_str());
}

float stof_strict(const std::string &s) { // xml.cpp:38-52
    size_t off = 0;
    float r = std::stof(s, &off);
    for (size_t i = off; i < s.size(); ++i)
        if (!isspace((unsigned char) s[i])) throw std::invalid_argument("trailing characters");
    return r;
}
int64_t stoll_strict(const std::string &s) {
    size_t off = 0;
    int64_t r = std::stoll(s, &off);
    for (size_t i = off; i < s.size(); ++i)
        if (!isspace((unsigned char) s[i])) throw std::invalid_argument("trailing characters");
    return r;
}

void check_attributes(const Context &ctx, const Node &n, std::set<std::string> attrs, bool expect_all = true) { // :192-208
    bool found_one = false;
    for (auto &a : n.attrs) {
        auto it = attrs.find(a.first);
        if (it == attrs.end()) node_error(ctx, n, format("unexpected attribute \"%s\" in element \"%s\"", a.first.c_str(), n.name.c_str()));
        attrs.erase(it);
        found_one = true;
    }
    if (!attrs.empty() && (!found_one || expect_all))
        node_error(ctx, n, format("missing attribute \"%s\" in element \"%s\"", attrs.begin()->c_str(), n.name.c_str()));
}

void expand_value_to_xyz(const Context &ctx, Node &n) { // :210-230
    if (!n.attr("value")) return;
    auto list = string::tokenize(n.value("value"));
    if (n.attr("x") || n.attr("y") || n.attr("z")) node_error(ctx, n, "can't mix and match \"value\" and \"x\"/\"y\"/\"z\" attributes");
    if (list.size() == 1) { n.append("x", list[0]); n.append("y", list[0]); n.append("z", list[0]); }
    else if (list.size() == 3) { n.append("x", list[0]); n.append("y", list[1]); n.append("z", list[2]); }
    else node_error(ctx, n, "\"value\" attribute must have exactly 1 or 3 elements");
    n.remove("value");
}

Vector3f parse_named_vector(const Context &ctx, const Node &n, const char *attr) { // :232-246
    std::string v = n.value(attr);
    auto list = string::tokenize(v);
    if (list.size() != 3) node_error(ctx, n, format("\"%s\" attribute must have exactly 3 elements", attr));
    try {
        return Vector3f{ stof_strict(list[0]), stof_strict(list[1]), stof_strict(list[2]) };
    } catch (...) {
        node_error(ctx, n, format("could not parse floating point values in \"%s\"", v.c_str()));
    }
}

Vector3f parse_vector(const Context &ctx, const Node &n, float def = 0.f) { // :248-267
    std::string value;
    try {
        Vector3f r{ def, def, def };
        value = n.value("x"); if (!value.empty()) r.x = stof_strict(value);
        value = n.value("y"); if (!value.empty()) r.y = stof_strict(value);
        value = n.value("z"); if (!value.empty()) r.z = stof_strict(value);
        return r;
    } catch (...) {
        node_error(ctx, n, format("could not parse floating point value \"%s\"", value.c_str()));
    }
}

ref<Object> create_texture_from_rgb(Color3 color, bool within_emitter) { // :269-277
    Properties p(within_emitter ? "srgb_d65" : "srgb");
    p.set_color("color", color);
    return InstanceManager::get()->create_instance(p, Class::for_name("Texture"));
}

ref<Object> create_texture_from_spectrum(float const_value, std::vector<float> &wavelengths, std::vector<float> &values,
                                         bool within_emitter) { // :279-342
    const Class *cls = Class::for_name("Texture");
    if (wavelengths.empty()) {
        Properties p("uniform");
        if (within_emitter) { p.set_instance_name("d65"); p.set_float("scale", const_value); }
        else p.set_float("value", const_value);
        ref<Object> obj = InstanceManager::get()->create_instance(p, cls);
        auto expanded = obj->expand();
        if (!expanded.empty()) obj = expanded[0];
        return obj;
    }
    const float cie_y_normalization = float(1.0 / 106.7502593994140625); // spectrum.h:75
    float unit_conversion = within_emitter ? cie_y_normalization : 1.f;
    bool is_regular = true;
    float interval = 0.f;
    for (size_t i = 0; i < wavelengths.size(); ++i) {
        values[i] *= unit_conversion;
        if (i == 0) continue;
        float distance = wavelengths[i] - wavelengths[i - 1];
        if (distance < 0.f) Throw("Wavelengths must be specified in increasing order!");
        if (i == 1) interval = distance;
        else if (std::abs(distance - interval) > 5.9604644775390625e-08f) is_regular = false;
    }
    if (!is_regular) Throw("Plugin \"irregular\" not found!"); // not compiled in the reference either (CMakeLists.txt:106-112)
    Properties p("regular");
    p.set_int("size", (int64_t) wavelengths.size());
    p.set_float("lambda_min", wavelengths.front());
    p.set_float("lambda_max", wavelengths.back());
    p.set_pointer("values", values.data());
    return InstanceManager::get()->create_instance(p, cls);
}

std::pair<std::string, std::string> parse_xml(Context &ctx, Node &node, Tag parent_tag, Properties &props, ParameterList &param,
                                              size_t &arg_counter, int depth, bool within_emitter = false, bool within_spectrum = false) {
    if (node.kind != Node::Element) return { "", "" };
    if (!param.empty())
        for (auto &a : node.attrs) {
            if (a.second.find('$') == std::string::npos) continue;
            for (auto &kv : param) {
                std::string key = "$" + kv.first;
                size_t p = 0;
                while ((p = a.second.find(key, p)) != std::string::npos) { a.second.replace(p, key.size(), kv.second); p += kv.second.size(); }
            }
        }
    auto it = tags().find(node.name);
    if (it == tags().end()) node_error(ctx, node, format("unexpected tag \"%s\"", node.name.c_str()));
    Tag tag = it->second;
    bool has_parent = parent_tag != Tag::Invalid, parent_is_object = has_parent && parent_tag == Tag::Object,
         current_is_object = tag == Tag::Object, parent_is_transform = parent_tag == Tag::Transform,
         current_is_transform_op = tag == Tag::Translate || tag == Tag::Rotate || tag == Tag::Scale || tag == Tag::LookAt || tag == Tag::Matrix;
    if (!has_parent && !current_is_object) node_error(ctx, node, format("root element \"%s\" must be an object", node.name.c_str()));
    if (parent_is_transform != current_is_transform_op) {
        if (parent_is_transform) node_error(ctx, node, "transform nodes can only contain transform operations");
        else node_error(ctx, node, "transform operations can only occur in a transform node");
    }
    if (has_parent && !parent_is_object && !(parent_is_transform && current_is_transform_op))
        node_error(ctx, node, format("node \"%s\" cannot occur as child of a property", node.name.c_str()));
    if (node.name == "scene") { if (!node.attr("type")) node.append("type", "scene"); }
    else if (tag == Tag::Transform) ctx.transform = Transform4f();
    if (node.attr("name")) {
        if (string::starts_with(node.value("name"), "_"))
            node_error(ctx, node, format("invalid parameter name \"%s\" in element \"%s\": leading underscores are reserved for internal identifiers.",
                                         node.value("name").c_str(), node.name.c_str()));
    } else if (current_is_object || tag == Tag::NamedReference) {
        node.append("name", format("_arg_%zu", arg_counter++));
    }
    if (node.attr("id")) {
        if (string::starts_with(node.value("id"), "_"))
            node_error(ctx, node, format("invalid id \"%s\" in element \"%s\": leading underscores are reserved for internal identifiers.",
                                         node.value("id").c_str(), node.name.c_str()));
    } else if (current_is_object) {
        node.append("id", format("_unnamed_%zu", ctx.id_counter++));
    }
    try {
        switch (tag) {
            case Tag::Object: {
                check_attributes(ctx, node, { "type", "id", "name" });
                std::string id = node.value("id"), name = node.value("name"), type = node.value("type");
                Properties nested(type);
                nested.set_id(id);
                auto prev = ctx.instances.find(id);
                if (prev != ctx.instances.end())
                    node_error(ctx, node, format("\"%s\" has duplicate id \"%s\" (previous was at line %d)", node.name.c_str(), id.c_str(), prev->second.line));
                auto cls = tag_class().find(node.name);
                if (cls == tag_class().end()) node_error(ctx, node, format("could not retrieve class object for tag \"%s\"", node.name.c_str()));
                size_t nested_counter = 0;
                for (Node &ch : node.children) {
                    auto [arg_name, nested_id] = parse_xml(ctx, ch, tag, nested, param, nested_counter, depth + 1, node.name == "emitter",
                                                           node.name == "spectrum");
                    if (!nested_id.empty()) nested.set_named_reference(arg_name, nested_id);
                }
                Instance &inst = ctx.instances[id];
                inst.props = nested;
                inst.clazz = cls->second;
                inst.line = node.line;
                return { name, id };
            }
            case Tag::NamedReference: {
                check_attributes(ctx, node, { "name", "id" });
                return { node.value("name"), node.value("id") };
            }
            case Tag::String:
                check_attributes(ctx, node, { "name", "value" });
                props.set_string(node.value("name"), node.value("value"));
                break;
            case Tag::Float: {
                check_attributes(ctx, node, { "name", "value" });
                float v;
                try { v = stof_strict(node.value("value")); }
                catch (...) { node_error(ctx, node, format("could not parse floating point value \"%s\"", node.value("value").c_str())); }
                props.set_float(node.value("name"), v);
                break;
            }
            case Tag::Integer: {
                check_attributes(ctx, node, { "name", "value" });
                int64_t v;
                try { v = stoll_strict(node.value("value")); }
                catch (...) { node_error(ctx, node, format("could not parse integer value \"%s\"", node.value("value").c_str())); }
                props.set_int(node.value("name"), v);
                break;
            }
            case Tag::Vector:
                expand_value_to_xyz(ctx, node);
                check_attributes(ctx, node, { "name", "x", "y", "z" });
                props.set_vector3(node.value("name"), parse_vector(ctx, node));
                break;
            case Tag::Matrix: {
                check_attributes(ctx, node, { "value" });
                auto tokens = string::tokenize(node.value("value"), " ");
                if (tokens.size() != 16) Throw("matrix: expected 16 values");
                float m[16];
                for (int i = 0; i < 16; ++i) {
                    try { m[i] = stof_strict(tokens[i]); }
                    catch (...) { node_error(ctx, node, format("could not parse floating point value \"%s\"", tokens[i].c_str())); }
                }
                ctx.transform = Transform4f(m) * ctx.transform;
                break;
            }
            case Tag::RGB: {
                check_attributes(ctx, node, { "name", "value" });
                auto tokens = string::tokenize(node.value("value"));
                if (tokens.size() == 1) { tokens.push_back(tokens[0]); tokens.push_back(tokens[0]); }
                if (tokens.size() != 3) node_error(ctx, node, format("'rgb' tag requires one or three values (got \"%s\")", node.value("value").c_str()));
                Color3 c;
                try { c = Color3{ stof_strict(tokens[0]), stof_strict(tokens[1]), stof_strict(tokens[2]) }; }
                catch (...) { node_error(ctx, node, format("could not parse RGB value \"%s\"", node.value("value").c_str())); }
                if (!within_spectrum) props.set_object(node.value("name"), create_texture_from_rgb(c, within_emitter));
                else props.set_color("color", c);
                break;
            }
            case Tag::Spectrum: {
                check_attributes(ctx, node, { "name", "value", "filename" }, false);
                std::vector<float> wavelengths, values;
                bool has_value = node.attr("value") && !node.value("value").empty(), has_filename = node.attr("filename") && !node.value("filename").empty();
                bool is_constant = has_value && string::tokenize(node.value("value")).size() == 1;
                float const_value = 1.f;
                if (has_value == has_filename) node_error(ctx, node, "'spectrum' tag requires one of \"value\" or \"filename\" attributes");
                else if (is_constant) {
                    try { const_value = stof_strict(string::tokenize(node.value("value"))[0]); }
                    catch (...) { node_error(ctx, node, format("could not parse constant spectrum \"%s\"", node.value("value").c_str())); }
                } else if (has_value) {
                    for (auto &token : string::tokenize(node.value("value"))) {
                        auto pair = string::tokenize(token, ":");
                        if (pair.size() != 2) node_error(ctx, node, "invalid spectrum (expected wavelength:value pairs)");
                        try { wavelengths.push_back(stof_strict(pair[0])); values.push_back(stof_strict(pair[1])); }
                        catch (...) { node_error(ctx, node, format("could not parse wavelength:value pair: \"%s\"", token.c_str())); }
                    }
                } else {
                    node_error(ctx, node, "'spectrum' tag: loading spectra from files is not implemented (nor in the reference, xml.cpp:620-622)");
                }
                props.set_object(node.value("name"), create_texture_from_spectrum(const_value, wavelengths, values, within_emitter));
                break;
            }
            case Tag::Transform:
                check_attributes(ctx, node, { "name" });
                ctx.transform = Transform4f();
                break;
            case Tag::LookAt: {
                check_attributes(ctx, node, { "origin", "target", "up" });
                Transform4f r = Transform4f::lookat(parse_named_vector(ctx, node, "origin"), parse_named_vector(ctx, node, "target"),
                                                    parse_named_vector(ctx, node, "up"));
                if (r.has_nan()) node_error(ctx, node, "invalid lookat transformation");
                ctx.transform = r * ctx.transform;
                break;
            }
            case Tag::Translate:
                expand_value_to_xyz(ctx, node);
                check_attributes(ctx, node, { "x", "y", "z" }, false);
                ctx.transform = Transform4f::translate(parse_vector(ctx, node)) * ctx.transform;
                break;
            case Tag::Scale:
                expand_value_to_xyz(ctx, node);
                check_attributes(ctx, node, { "x", "y", "z" }, false);
                ctx.transform = Transform4f::scale(parse_vector(ctx, node, 1.f)) * ctx.transform;
                break;
            case Tag::Boolean: {
                check_attributes(ctx, node, { "name", "value" });
                std::string v = string::to_lower(node.value("value"));
                if (v == "true") props.set_bool(node.value("name"), true);
                else if (v == "false") props.set_bool(node.value("name"), false);
                else node_error(ctx, node, format("could not parse boolean value \"%s\" -- must be \"true\" or \"false\"", node.value("value").c_str()));
                break;
            }
            case Tag::Rotate: {
                expand_value_to_xyz(ctx, node);
                check_attributes(ctx, node, { "angle", "x", "y", "z" }, false);
                if (!node.attr("angle")) node_error(ctx, node, "missing attribute \"angle\" in element \"rotate\"");
                Vector3f axis = parse_vector(ctx, node);
                if (axis.x == 0.f && axis.y == 0.f && axis.z == 0.f) node_error(ctx, node, "rotate: the axis must not be zero");
                float angle;
                try { angle = stof_strict(node.value("angle")); }
                catch (...) { node_error(ctx, node, format("could not parse floating point value \"%s\"", node.value("angle").c_str())); }
                ctx.transform = Transform4f::rotate(axis, angle * (3.14159265358979323846f / 180.f)) * ctx.transform;
                break;
            }
            case Tag::Default: {
                check_attributes(ctx, node, { "name", "value" });
                std::string name = node.value("name");
                if (name.empty()) node_error(ctx, node, "<default>: name must be nonempty");
                bool found = false;
                for (auto &kv : param) found |= kv.first == name;
                if (!found) param.emplace_back(name, node.value("value"));
                break;
            }
            case Tag::Alias: {
                check_attributes(ctx, node, { "id", "as" });
                std::string id = node.value("id"), as = node.value("as");
                if (ctx.instances.find(id) == ctx.instances.end() && ctx.aliases.find(id) == ctx.aliases.end())
                    node_error(ctx, node, format("referenced id \"%s\" not found", id.c_str()));
                if (ctx.instances.find(as) != ctx.instances.end() || ctx.aliases.find(as) != ctx.aliases.end())
                    node_error(ctx, node, format("duplicate id \"%s\"", as.c_str()));
                auto chained = ctx.aliases.find(id);
                ctx.aliases[as] = chained != ctx.aliases.end() ? chained->second : id;
                break;
            }
            case Tag::Include: {
                check_attributes(ctx, node, { "filename" });
                std::string filename = node.value("filename");
                if (filename.empty() || filename[0] != '/') {
                    size_t slash = ctx.src.find_last_of('/');
                    if (slash != std::string::npos) filename = ctx.src.substr(0, slash + 1) + filename;
                }
                std::ifstream is(filename, std::ios::binary);
                if (!is) { // not next to the including file: the FileResolver's search path (fresolver.cpp)
                    filename = get_file_resolver()->resolve(node.value("filename"));
                    is.open(filename, std::ios::binary);
                }
                if (!is) node_error(ctx, node, format("included file \"%s\" not found", filename.c_str()));
                if (ctx.include_depth >= 15) node_error(ctx, node, "exceeded <include> recursion limit of 15");
                std::stringstream ss;
                ss << is.rdbuf();
                const std::string text = ss.str(); // the Reader keeps references to both strings
                Reader reader(text, filename);
                Node root = reader.parse_document();
                std::string outer_src = ctx.src;
                ctx.src = filename;
                ++ctx.include_depth;
                auto attach = [&](Node &ch) {
                    auto [arg_name, nested_id] = parse_xml(ctx, ch, parent_tag, props, param, arg_counter, depth, within_emitter, within_spectrum);
                    if (!nested_id.empty()) props.set_named_reference(arg_name, nested_id);
                };
                if (root.name == "scene") for (Node &ch : root.children) attach(ch);
                else attach(root);
                --ctx.include_depth;
                ctx.src = outer_src;
                break;
            }
            default: break;
        }
        for (Node &ch : node.children) parse_xml(ctx, ch, tag, props, param, arg_counter, depth + 1);
        if (tag == Tag::Transform) props.set_transform(node.value("name"), ctx.transform);
    } catch (const std::exception &e) {
        if (strstr(e.what(), "Error while loading") == nullptr) node_error(ctx, node, e.what());
        throw;
    }
    return { "", "" };
}

ref<Object> instantiate_node(Context &ctx, const std::string &id) { // :676-710
    auto alias = ctx.aliases.find(id);
    auto it = ctx.instances.find(alias != ctx.aliases.end() ? alias->second : id);
    if (it == ctx.instances.end()) Throw("reference to unknown object \"%s\"!", id.c_str());
    Instance &inst = it->second;
    if (inst.object) return inst.object;
    for (auto &kv : inst.props.named_references()) {
        try {
            inst.props.set_object(kv.first, instantiate_node(ctx, kv.second));
        } catch (const std::exception &e) {
            if (strstr(e.what(), "Error while loading") == nullptr)
                Throw("Error while loading \"%s\" (near line %d): %s", ctx.src.c_str(), inst.line, e.what());
            throw;
        }
    }
    try {
        inst.object = InstanceManager::get()->create_instance(inst.props, inst.clazz);
    } catch (const std::exception &e) {
        Throw("Error while loading \"%s\" (near line %d): could not instantiate %s instance of type \"%s\": %s", ctx.src.c_str(), inst.line,
              string::to_lower(inst.clazz->name()).c_str(), inst.props.instance_name().c_str(), e.what());
    }
    return inst.object;
}

} // namespace

void register_class(const Class *c) { // xml.cpp:65-101
    const std::string &alias = c->alias();
    if (tags().find(alias) == tags().end()) tags()[alias] = Tag::Object;
    tag_class()[alias] = c;
    if (alias == "texture") tag_class()["spectrum"] = c;
}

ref<Object> load_string(const std::string &text, const std::string &source_id, ParameterList parameters) {
    Class::static_initialization();
    Reader reader(text, source_id);
    Node root = reader.parse_document();
    Context ctx;
    ctx.src = source_id;
    Properties props;
    size_t arg_counter = 0;
    auto [name, id] = parse_xml(ctx, root, Tag::Invalid, props, parameters, arg_counter, 0);
    return instantiate_node(ctx, id);
}

ref<Object> load_file(const std::string &filename, ParameterList parameters) {
    std::ifstream is(filename, std::ios::binary);
    if (!is) Throw("\"%s\": file not exists.", filename.c_str());
    Log(Info, "Loading XML file \"%s\" ..", filename.c_str());
    std::stringstream ss;
    ss << is.rdbuf();
    return load_string(ss.str(), filename, std::move(parameters));
}

} // namespace xml
} // namespace misaki
