// scene_io.cpp -- host-side scene I/O of the backend (C++17; the reference's host code is Rust, which this
// image cannot compile, so the layer above the C ABI is C++).
//
// Mirrors: SceneDescriptor::load + to_data (src/render/mod.rs:92-110), SceneObjectDescriptorType::to_scene_object
// (:304-318), Mesh::new (:450-499), load_off (src/render/load_off.rs:8-85), gamma (:57-63), the PPM writer
// (:1042-1076) and hash_vec_of_vectors (:916-926).  Formats are kept verbatim: serde_json externally tagged enums,
// floats parsed as f64 then narrowed to f32; OFF numbers parsed straight to f32.
#include "scene_io.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

namespace ptb {

// ------------------------------------------------------------------------------------------------
// a small JSON document model
// ------------------------------------------------------------------------------------------------
struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    double number = 0.0;
    bool boolean = false;
    std::string text;
    std::vector<Json> elems;                           // Array
    std::vector<std::pair<std::string, Json>> members;  // Object, insertion order

    const Json *get(const char *key) const {
        if (kind != Object) return nullptr;
        for (const auto &m : members)
            if (m.first == key) return &m.second;
        return nullptr;
    }
};

class JsonReader {
  public:
    explicit JsonReader(const std::string &s) : src_(s) {}
    Json parse() {
        Json v = value();
        skip();
        if (pos_ != src_.size()) fail("trailing characters");
        return v;
    }

  private:
    const std::string &src_;
    size_t pos_ = 0;

    [[noreturn]] void fail(const std::string &what) const {
        size_t line = 1, col = 1;
        for (size_t i = 0; i < pos_ && i < src_.size(); ++i) {
            if (src_[i] == '\n') { ++line; col = 1; } else ++col;
        }
        throw SceneError(PTB_ERR_PARSE, what + " at line " + std::to_string(line) + " column " + std::to_string(col));
    }
    void skip() {
        while (pos_ < src_.size() && std::isspace(static_cast<unsigned char>(src_[pos_]))) ++pos_;
    }
    bool eat(char c) {
        skip();
        if (pos_ < src_.size() && src_[pos_] == c) { ++pos_; return true; }
        return false;
    }
    std::string string_lit() {
        if (!eat('"')) fail("expected a string");
        std::string out;
        while (true) {
            if (pos_ >= src_.size()) fail("EOF while parsing a string");
            char c = src_[pos_++];
            if (c == '"') break;
            if (c == '\\') {
                if (pos_ >= src_.size()) fail("EOF in escape");
                char e = src_[pos_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (pos_ + 4 > src_.size()) fail("bad \\u escape");
                        unsigned cp = static_cast<unsigned>(std::strtoul(src_.substr(pos_, 4).c_str(), nullptr, 16));
                        pos_ += 4;
                        if (cp < 0x80) out += static_cast<char>(cp);
                        else if (cp < 0x800) { out += static_cast<char>(0xC0 | (cp >> 6)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                        else { out += static_cast<char>(0xE0 | (cp >> 12)); out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                    } break;
                    default: out += e;
                }
            } else out += c;
        }
        return out;
    }
    Json value() {
        skip();
        if (pos_ >= src_.size()) fail("EOF while parsing a value");
        Json v;
        char c = src_[pos_];
        if (c == '{') {
            ++pos_;
            v.kind = Json::Object;
            if (eat('}')) return v;
            do {
                std::string key = string_lit();
                if (!eat(':')) fail("expected `:`");
                v.members.emplace_back(std::move(key), value());
            } while (eat(','));
            if (!eat('}')) fail("expected `,` or `}`");
        } else if (c == '[') {
            ++pos_;
            v.kind = Json::Array;
            if (eat(']')) return v;
            do { v.elems.push_back(value()); } while (eat(','));
            if (!eat(']')) fail("expected `,` or `]`");
        } else if (c == '"') {
            v.kind = Json::String;
            v.text = string_lit();
        } else if (src_.compare(pos_, 4, "null") == 0) {
            pos_ += 4;
        } else if (src_.compare(pos_, 4, "true") == 0) {
            pos_ += 4; v.kind = Json::Bool; v.boolean = true;
        } else if (src_.compare(pos_, 5, "false") == 0) {
            pos_ += 5; v.kind = Json::Bool;
        } else {
            const char *b = src_.c_str() + pos_;
            char *e = nullptr;
            double dv = std::strtod(b, &e);
            if (e == b) fail("expected value");
            pos_ += static_cast<size_t>(e - b);
            v.kind = Json::Number;
            v.number = dv;
        }
        return v;
    }
};

static float need_f32(const Json *j, const char *what) {
    if (!j) throw SceneError(PTB_ERR_PARSE, std::string("missing field `") + what + "`");
    if (j->kind != Json::Number) throw SceneError(PTB_ERR_PARSE, std::string("invalid type for `") + what + "`: expected f32");
    return static_cast<float>(j->number);  // serde: f64 -> `as f32`
}
static void need_vec3(const Json *j, const char *what, float out[3]) {
    if (!j) throw SceneError(PTB_ERR_PARSE, std::string("missing field `") + what + "`");
    if (j->kind != Json::Array || j->elems.size() != 3)
        throw SceneError(PTB_ERR_PARSE, std::string("invalid `") + what + "`: expected an array of 3 numbers");
    for (int i = 0; i < 3; ++i) out[i] = need_f32(&j->elems[static_cast<size_t>(i)], what);
}

// ------------------------------------------------------------------------------------------------
// Mesh::new (mod.rs:450-499): the bounding sphere the gate test uses; centre = min + max*0.5 (sic)
// ------------------------------------------------------------------------------------------------
void mesh_bounding_sphere(const ptb_triangle *tris, size_t n, float centre[3], float *radius) {
    const float inf = std::numeric_limits<float>::infinity();
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (size_t i = 0; i < n; ++i) {
        const float *verts[3] = {tris[i].a, tris[i].b, tris[i].c};
        for (const float *v : verts)
            for (int k = 0; k < 3; ++k) {
                if (v[k] < lo[k]) lo[k] = v[k];
                if (v[k] > hi[k]) hi[k] = v[k];
            }
    }
    for (int k = 0; k < 3; ++k) centre[k] = lo[k] + hi[k] * 0.5f;
    auto dist = [&](const float p[3]) {
        float dx = p[0] - centre[0], dy = p[1] - centre[1], dz = p[2] - centre[2];
        return std::sqrt((dx * dx + dy * dy) + dz * dz);
    };
    float r_lo = dist(lo), r_hi = dist(hi);
    *radius = r_hi >= r_lo ? r_hi : r_lo;
}

// ------------------------------------------------------------------------------------------------
// load_off (load_off.rs:8-85)
// ------------------------------------------------------------------------------------------------
namespace {
struct OffLines {
    std::ifstream in;
    // next non-empty, non-'#' line, trimmed (load_off.rs:12-20)
    bool next(std::string &line) {
        std::string raw;
        while (std::getline(in, raw)) {
            size_t b = 0, e = raw.size();
            while (b < e && std::isspace(static_cast<unsigned char>(raw[b]))) ++b;
            while (e > b && std::isspace(static_cast<unsigned char>(raw[e - 1]))) --e;
            if (e == b || raw[b] == '#') continue;
            line.assign(raw, b, e - b);
            return true;
        }
        return false;
    }
};
std::vector<std::string> split_ws(const std::string &s) {
    std::vector<std::string> out;
    std::istringstream ss(s);
    std::string tok;
    while (ss >> tok) out.push_back(tok);
    return out;
}
bool parse_usize(const std::string &tok, size_t &out) {  // str::parse::<usize>
    size_t i = 0;
    if (!tok.empty() && tok[0] == '+') i = 1;
    if (i >= tok.size()) return false;
    size_t v = 0;
    for (; i < tok.size(); ++i) {
        if (tok[i] < '0' || tok[i] > '9') return false;
        v = v * 10 + static_cast<size_t>(tok[i] - '0');
    }
    out = v;
    return true;
}
bool parse_f32(const std::string &tok, float &out) {  // str::parse::<f32>: correctly rounded, whole token
    const char *b = tok.c_str();
    char *e = nullptr;
    out = std::strtof(b, &e);
    return e != b && *e == '\0';
}
}  // namespace

void load_off(const std::string &path, float scale, std::vector<ptb_triangle> &out, unsigned flags) {
    OffLines L;
    L.in.open(path);
    if (!L.in) throw SceneError(PTB_ERR_IO, "cannot open " + path);
    auto bad = [&](const std::string &why) -> SceneError { return SceneError(PTB_ERR_PARSE, path + ": " + why); };
    std::string line;
    if (!L.next(line) || line != "OFF") throw bad("Invalid header");
    if (!L.next(line)) throw bad("Invalid element counts");
    auto toks = split_ws(line);
    size_t counts[3];
    if (toks.size() != 3) throw bad("Invalid element counts");
    for (int i = 0; i < 3; ++i)
        if (!parse_usize(toks[static_cast<size_t>(i)], counts[i])) throw bad("Invalid element counts");
    const size_t nv = counts[0], nf = counts[1];
    std::vector<float> verts(nv * 3);
    for (size_t i = 0; i < nv; ++i) {
        if (!L.next(line)) throw bad("unexpected end of file in vertex list");
        toks = split_ws(line);
        float c[3];
        if (toks.size() != 3 || !parse_f32(toks[0], c[0]) || !parse_f32(toks[1], c[1]) || !parse_f32(toks[2], c[2]))
            throw bad("Invalid vertex coordinates");
        for (int k = 0; k < 3; ++k) verts[i * 3 + static_cast<size_t>(k)] = c[k] * scale;  // load_off.rs:52
    }
    out.reserve(out.size() + nf);
    for (size_t i = 0; i < nf; ++i) {
        if (!L.next(line)) throw bad("unexpected end of file in face list");
        toks = split_ws(line);
        size_t idx[4];
        bool ok = toks.size() >= 4;
        for (int k = 0; ok && k < 4; ++k) ok = parse_usize(toks[static_cast<size_t>(k)], idx[k]);
        auto emit = [&](size_t ia, size_t ib, size_t ic) {
            ptb_triangle t;
            for (int k = 0; k < 3; ++k) {
                t.a[k] = verts[ia * 3 + static_cast<size_t>(k)];
                t.b[k] = verts[ib * 3 + static_cast<size_t>(k)];
                t.c[k] = verts[ic * 3 + static_cast<size_t>(k)];
            }
            out.push_back(t);
        };
        if (ok && idx[0] > 3 && (flags & PTB_LOAD_TRIANGULATE_POLYGONS)) {
            // NOT reference behaviour (opt-in): fan-triangulate an n-gon (v0, v_i, v_i+1); the reference rejects it
            const size_t cnt = idx[0];
            std::vector<size_t> poly(cnt);
            bool good = toks.size() >= cnt + 1;
            for (size_t k = 0; good && k < cnt; ++k) good = parse_usize(toks[k + 1], poly[k]) && poly[k] < nv;
            if (!good) throw bad("Invalid face: " + line);
            for (size_t k = 1; k + 1 < cnt; ++k) emit(poly[0], poly[k], poly[k + 1]);
            continue;
        }
        // only triangles are supported (load_off.rs:73-76); trailing colour tokens are ignored
        if (!ok || idx[0] != 3 || idx[1] >= nv || idx[2] >= nv || idx[3] >= nv) throw bad("Invalid face: " + line);
        emit(idx[1], idx[2], idx[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// scenes/<id>.json -> HostScene
// ------------------------------------------------------------------------------------------------
static int parse_reflect_type(const Json *j) {
    if (!j) throw SceneError(PTB_ERR_PARSE, "missing field `reflect_type`");
    if (j->kind != Json::String) throw SceneError(PTB_ERR_PARSE, "invalid type for `reflect_type`");
    if (j->text == "Diffuse") return PTB_REFL_DIFFUSE;
    if (j->text == "Specular") return PTB_REFL_SPECULAR;
    if (j->text == "Refract") return PTB_REFL_REFRACT;
    throw SceneError(PTB_ERR_PARSE, "unknown variant `" + j->text + "`, expected one of `Diffuse`, `Specular`, `Refract`");
}

HostScene load_scene_json(const std::string &json_path, const std::string &base_dir, unsigned flags) {
    std::ifstream in(json_path, std::ios::binary);
    if (!in) throw SceneError(PTB_ERR_IO, "cannot open " + json_path);
    std::stringstream buf;
    buf << in.rdbuf();
    const std::string text = buf.str();
    const Json root = JsonReader(text).parse();
    if (root.kind != Json::Object) throw SceneError(PTB_ERR_PARSE, "scene file is not a JSON object");

    HostScene sc;
    const Json *id = root.get("id");
    if (!id || id->kind != Json::String) throw SceneError(PTB_ERR_PARSE, "missing field `id`");
    sc.id = id->text;

    const Json *cam = root.get("camera");
    if (!cam || cam->kind != Json::Object) throw SceneError(PTB_ERR_PARSE, "missing field `camera`");
    need_vec3(cam->get("position"), "position", sc.camera.position);
    need_vec3(cam->get("direction"), "direction", sc.camera.direction);
    sc.camera.focal_length = need_f32(cam->get("focal_length"), "focal_length");
    sc.camera.sensor_width = need_f32(cam->get("sensor_width"), "sensor_width");
    sc.camera.aspect_ratio = need_f32(cam->get("aspect_ratio"), "aspect_ratio");

    const Json *objs = root.get("objects");
    if (!objs || objs->kind != Json::Array) throw SceneError(PTB_ERR_PARSE, "missing field `objects`");
    for (size_t i = 0; i < objs->elems.size(); ++i) {
        const Json &jo = objs->elems[i];
        try {
            ptb_object o;
            std::memset(&o, 0, sizeof o);
            need_vec3(jo.get("position"), "position", o.position);
            const Json *mat = jo.get("material");
            if (!mat) throw SceneError(PTB_ERR_PARSE, "missing field `material`");
            need_vec3(mat->get("color"), "color", o.color);
            need_vec3(mat->get("emmission"), "emmission", o.emission);
            o.reflect_type = parse_reflect_type(mat->get("reflect_type"));
            const Json *ty = jo.get("type_");
            if (!ty || ty->kind != Json::Object || ty->members.size() != 1)
                throw SceneError(PTB_ERR_PARSE, "invalid `type_`: expected a single-variant map");
            const std::string &variant = ty->members[0].first;
            const Json &body = ty->members[0].second;
            o.tri_begin = sc.triangles.size();
            ObjectSource src;
            if (variant == "Sphere") {
                o.kind = PTB_OBJ_SPHERE;
                o.radius = need_f32(body.get("radius"), "radius");
            } else if (variant == "MeshFile") {
                o.kind = PTB_OBJ_MESH;
                const Json *p = body.get("path");
                if (!p || p->kind != Json::String) throw SceneError(PTB_ERR_PARSE, "missing field `path`");
                const float scale = need_f32(body.get("scale"), "scale");
                std::string full = p->text;
                if (!full.empty() && full[0] != '/' && !base_dir.empty()) full = base_dir + "/" + full;
                src.variant = 1; src.path = p->text; src.scale = scale;
                load_off(full, scale, sc.triangles, flags);
                o.tri_count = sc.triangles.size() - o.tri_begin;
                mesh_bounding_sphere(sc.triangles.data() + o.tri_begin, o.tri_count, o.bs_position, &o.bs_radius);
            } else if (variant == "Mesh") {
                o.kind = PTB_OBJ_MESH;
                const Json *tris = body.get("triangles");
                if (!tris || tris->kind != Json::Array) throw SceneError(PTB_ERR_PARSE, "missing field `triangles`");
                for (const Json &jt : tris->elems) {
                    ptb_triangle t;
                    need_vec3(jt.get("a"), "a", t.a);
                    need_vec3(jt.get("b"), "b", t.b);
                    need_vec3(jt.get("c"), "c", t.c);
                    sc.triangles.push_back(t);
                }
                o.tri_count = sc.triangles.size() - o.tri_begin;
                // deserialised, not recomputed (mod.rs:441-448)
                const Json *bs = body.get("bounding_sphere");
                if (!bs) throw SceneError(PTB_ERR_PARSE, "missing field `bounding_sphere`");
                need_vec3(bs->get("position"), "position", o.bs_position);
                o.bs_radius = need_f32(bs->get("radius"), "radius");
                const Json *bb = body.get("bounding_box");
                if (!bb || bb->kind != Json::Array) throw SceneError(PTB_ERR_PARSE, "missing field `bounding_box`");
                src.variant = 2;
                for (const Json &jt : bb->elems) {
                    ptb_triangle t;
                    need_vec3(jt.get("a"), "a", t.a);
                    need_vec3(jt.get("b"), "b", t.b);
                    need_vec3(jt.get("c"), "c", t.c);
                    src.bounding_box.push_back(t);
                }
            } else {
                throw SceneError(PTB_ERR_PARSE, "unknown variant `" + variant + "`, expected one of `Sphere`, `MeshFile`, `Mesh`");
            }
            sc.objects.push_back(o);
            sc.sources.push_back(std::move(src));
        } catch (const SceneError &e) {
            throw SceneError(e.code, "objects[" + std::to_string(i) + "]: " + e.what());
        }
    }
    sc.refresh_desc();
    return sc;
}

// ------------------------------------------------------------------------------------------------
// write-back: SceneData::to_descriptor + SceneDescriptor::save (mod.rs:112-150)
// ------------------------------------------------------------------------------------------------
// serde_json prints an f32 with ryu: the shortest decimal that parses back to the same f32, plain notation while the decimal
// point stays within the digits' reach (ryu's pretty printer), exponent notation otherwise, always with a fraction ("1.0").
std::string format_f32(float v) {
    if (std::isnan(v) || std::isinf(v)) return "null";  // serde_json writes non-finite floats as null
    if (v == 0.0f) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 9; ++prec) {
        std::snprintf(buf, sizeof buf, "%.*e", prec - 1, static_cast<double>(v));
        if (std::strtof(buf, nullptr) == v) break;
    }
    std::string digits;
    bool neg = false;
    int exp10 = 0;
    {
        const char *p = buf;
        if (*p == '-') { neg = true; ++p; }
        for (; *p && *p != 'e'; ++p)
            if (*p != '.') digits += *p;
        exp10 = std::atoi(p + 1);
    }
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int len = static_cast<int>(digits.size());
    const int k = exp10 - (len - 1);  // value = digits * 10^k
    const int kk = len + k;           // position of the decimal point
    std::string out = neg ? "-" : "";
    if (0 <= k && kk <= 16) {
        out += digits + std::string(static_cast<size_t>(k), '0') + ".0";
    } else if (0 < kk && kk <= 16) {
        out += digits.substr(0, static_cast<size_t>(kk)) + "." + digits.substr(static_cast<size_t>(kk));
    } else if (-5 < kk && kk <= 0) {
        out += "0." + std::string(static_cast<size_t>(-kk), '0') + digits;
    } else {
        out += digits.substr(0, 1);
        if (len > 1) out += "." + digits.substr(1);
        out += "e" + std::to_string(kk - 1);
    }
    return out;
}

namespace {
struct PrettyWriter {
    std::string out;
    int depth = 0;
    void nl() { out += "\n"; out.append(static_cast<size_t>(depth) * 2, ' '); }
    void vec3(const float *v) {
        out += "[";
        ++depth;
        for (int i = 0; i < 3; ++i) { nl(); out += format_f32(v[i]); if (i < 2) out += ","; }
        --depth; nl();
        out += "]";
    }
    void str(const std::string &s) {
        out += '"';
        for (char c : s) {
            if (c == '"' || c == '\\') { out += '\\'; out += c; }
            else if (c == '\n') out += "\\n";
            else if (c == '\t') out += "\\t";
            else out += c;
        }
        out += '"';
    }
    void triangle(const ptb_triangle &t) {
        out += "{"; ++depth;
        nl(); out += "\"a\": "; vec3(t.a); out += ",";
        nl(); out += "\"b\": "; vec3(t.b); out += ",";
        nl(); out += "\"c\": "; vec3(t.c);
        --depth; nl(); out += "}";
    }
    void triangles(const ptb_triangle *t, size_t n) {
        if (n == 0) { out += "[]"; return; }
        out += "["; ++depth;
        for (size_t i = 0; i < n; ++i) { nl(); triangle(t[i]); if (i + 1 < n) out += ","; }
        --depth; nl(); out += "]";
    }
};
}  // namespace

std::string scene_to_json(const HostScene &sc) {
    static const char *kRefl[] = {"Diffuse", "Specular", "Refract"};
    PrettyWriter w;
    w.out += "{"; ++w.depth;
    w.nl(); w.out += "\"id\": "; w.str(sc.id); w.out += ",";
    w.nl(); w.out += "\"objects\": ";
    if (sc.objects.empty()) w.out += "[]";
    else {
        w.out += "["; ++w.depth;
        for (size_t i = 0; i < sc.objects.size(); ++i) {
            const ptb_object &o = sc.objects[i];
            const ObjectSource src = i < sc.sources.size() ? sc.sources[i] : ObjectSource{};
            w.nl(); w.out += "{"; ++w.depth;
            w.nl(); w.out += "\"type_\": {"; ++w.depth;
            if (o.kind == PTB_OBJ_SPHERE) {
                w.nl(); w.out += "\"Sphere\": {"; ++w.depth;
                w.nl(); w.out += "\"radius\": " + format_f32(o.radius);
                --w.depth; w.nl(); w.out += "}";
            } else if (src.variant == 1) {
                w.nl(); w.out += "\"MeshFile\": {"; ++w.depth;
                w.nl(); w.out += "\"path\": "; w.str(src.path); w.out += ",";
                w.nl(); w.out += "\"scale\": " + format_f32(src.scale);
                --w.depth; w.nl(); w.out += "}";
            } else {
                w.nl(); w.out += "\"Mesh\": {"; ++w.depth;
                w.nl(); w.out += "\"triangles\": "; w.triangles(sc.triangles.data() + o.tri_begin, o.tri_count); w.out += ",";
                w.nl(); w.out += "\"bounding_sphere\": {"; ++w.depth;
                w.nl(); w.out += "\"position\": "; w.vec3(o.bs_position); w.out += ",";
                w.nl(); w.out += "\"radius\": " + format_f32(o.bs_radius);
                --w.depth; w.nl(); w.out += "},";
                w.nl(); w.out += "\"bounding_box\": "; w.triangles(src.bounding_box.data(), src.bounding_box.size());
                --w.depth; w.nl(); w.out += "}";
            }
            --w.depth; w.nl(); w.out += "},";
            w.nl(); w.out += "\"position\": "; w.vec3(o.position); w.out += ",";
            w.nl(); w.out += "\"material\": {"; ++w.depth;
            w.nl(); w.out += "\"color\": "; w.vec3(o.color); w.out += ",";
            w.nl(); w.out += "\"emmission\": "; w.vec3(o.emission); w.out += ",";
            w.nl(); w.out += std::string("\"reflect_type\": \"") + kRefl[o.reflect_type] + "\"";
            --w.depth; w.nl(); w.out += "}";
            --w.depth; w.nl(); w.out += "}";
            if (i + 1 < sc.objects.size()) w.out += ",";
        }
        --w.depth; w.nl(); w.out += "]";
    }
    w.out += ",";
    w.nl(); w.out += "\"camera\": {"; ++w.depth;
    w.nl(); w.out += "\"position\": "; w.vec3(sc.camera.position); w.out += ",";
    w.nl(); w.out += "\"direction\": "; w.vec3(sc.camera.direction); w.out += ",";
    w.nl(); w.out += "\"focal_length\": " + format_f32(sc.camera.focal_length) + ",";
    w.nl(); w.out += "\"sensor_width\": " + format_f32(sc.camera.sensor_width) + ",";
    w.nl(); w.out += "\"aspect_ratio\": " + format_f32(sc.camera.aspect_ratio);
    --w.depth; w.nl(); w.out += "}";
    --w.depth; w.nl(); w.out += "}";
    return w.out;
}

void save_scene_json(const HostScene &sc, const std::string &json_path) {
    std::ofstream f(json_path, std::ios::binary);
    if (!f) throw SceneError(PTB_ERR_IO, "cannot create " + json_path);
    const std::string text = scene_to_json(sc);
    f.write(text.data(), static_cast<std::streamsize>(text.size()));
    if (!f) throw SceneError(PTB_ERR_IO, "write error on " + json_path);
}

// ------------------------------------------------------------------------------------------------
// output resolve
// ------------------------------------------------------------------------------------------------
uint32_t to_int_with_gamma_correction(float x) {  // mod.rs:57-63
    float c = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return static_cast<uint32_t>(255.0f * std::pow(c, 1.0f / 2.2f) + 0.5f);
}

void write_ppm(const std::string &path, const float *mean_rgb, int W, int H, uint64_t spp, const std::string &scene_id,
               uint64_t seconds) {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw SceneError(PTB_ERR_IO, "cannot create " + path);
    std::fprintf(f, "P3\n# samplesPerPixel: %llu, resolution_y: %d, scene_id: %s\n", static_cast<unsigned long long>(spp), H,
                 scene_id.c_str());
    std::fprintf(f, "# rendering time: %llu s\n", static_cast<unsigned long long>(seconds));
    std::fprintf(f, "%d %d\n%d\n", W, H, 255);
    std::string out;
    out.reserve(static_cast<size_t>(W) * static_cast<size_t>(H) * 12);
    char tmp[48];
    for (long long i = static_cast<long long>(W) * H - 1; i >= 0; --i) {  // pixels.iter().rev(), mod.rs:1065
        int n = std::snprintf(tmp, sizeof tmp, "%u %u %u ", to_int_with_gamma_correction(mean_rgb[3 * i]),
                              to_int_with_gamma_correction(mean_rgb[3 * i + 1]), to_int_with_gamma_correction(mean_rgb[3 * i + 2]));
        out.append(tmp, static_cast<size_t>(n));
    }
    std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
}

// hash_vec_of_vectors (mod.rs:916-926): std DefaultHasher = SipHash-1-3 with a zero key over the little-endian
// bytes of every component's bit pattern, fed as u32 writes.
namespace {
inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
struct Sip13 {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL, v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    uint64_t tail = 0;
    unsigned ntail = 0;
    uint64_t length = 0;
    void round() {
        v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
        v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
    }
    void write_u32(uint32_t x) {
        length += 4;
        tail |= static_cast<uint64_t>(x) << (8 * ntail);
        ntail += 4;
        if (ntail == 8) {
            v3 ^= tail; round(); v0 ^= tail;
            tail = 0; ntail = 0;
        }
    }
    uint64_t finish() {
        uint64_t b = ((length & 0xff) << 56) | tail;
        v3 ^= b; round(); v0 ^= b;
        v2 ^= 0xff;
        round(); round(); round();
        return v0 ^ v1 ^ v2 ^ v3;
    }
};
}  // namespace

uint64_t hash_pixels(const float *rgb, uint64_t n_pixels) {
    Sip13 h;
    for (uint64_t i = 0; i < n_pixels * 3; ++i) {
        uint32_t bits;
        std::memcpy(&bits, &rgb[i], 4);
        h.write_u32(bits);
    }
    return h.finish();
}

}  // namespace ptb
