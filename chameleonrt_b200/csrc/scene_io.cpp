// scene_io.cpp — the native scene loader behind include/crt_scene_io.h.
//
// Produces what the reference's Scene::load_obj produces (util/scene.cpp:94-228 over tinyobjloader 1.4.x, vendored by the
// reference as util/tiny_obj_loader.h; its parsing rules are restated here where the result depends on them, with the line
// they follow), from a memory-mapped file parsed by several threads. Host code only: no CUDA in this translation unit.
#include "../../include/crt_scene_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "host_parallel.h"
#include "image_decode.h"
#include "json_reader.h"

namespace {

thread_local std::string g_last_error;

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------------------------
// Numbers as tinyobjloader reads them (tiny_obj_loader.h:567-697): decimal digits accumulated in a double, fraction
// digits scaled by a table / pow(10, -k), the exponent applied as ldexp(m * pow(5, e), e), then rounded to float. This is
// not strtod (it is not correctly rounded), and the loader's output must have tinyobjloader's bits.
// ---------------------------------------------------------------------------------------------------------------
inline bool is_digit(char c)
{
    return (unsigned)(c - '0') < 10u;
}

bool try_parse_double(const char *s, const char *s_end, double *result)
{
    if (s >= s_end) {
        return false;
    }
    double mantissa = 0.0;
    int exponent = 0;
    char sign = '+', exp_sign = '+';
    const char *curr = s;
    int read = 0;
    bool end_not_reached = false;
    if (*curr == '+' || *curr == '-') {
        sign = *curr;
        curr++;
    } else if (!is_digit(*curr)) {
        return false;
    }
    end_not_reached = curr != s_end;
    while (end_not_reached && is_digit(*curr)) {
        mantissa *= 10;
        mantissa += (int)(*curr - '0');
        curr++;
        read++;
        end_not_reached = curr != s_end;
    }
    if (read == 0) {
        return false;
    }
    bool assemble = !end_not_reached;
    if (!assemble) {
        if (*curr == '.') {
            curr++;
            read = 1;
            end_not_reached = curr != s_end;
            static const double pow_lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            const int lut_entries = (int)(sizeof pow_lut / sizeof pow_lut[0]);
            while (end_not_reached && is_digit(*curr)) {
                mantissa += (int)(*curr - '0') * (read < lut_entries ? pow_lut[read] : std::pow(10.0, -read));
                read++;
                curr++;
                end_not_reached = curr != s_end;
            }
        } else if (*curr == 'e' || *curr == 'E') {
        } else {
            assemble = true;
        }
    }
    if (!assemble && end_not_reached && (*curr == 'e' || *curr == 'E')) {
        curr++;
        end_not_reached = curr != s_end;
        if (end_not_reached && (*curr == '+' || *curr == '-')) {
            exp_sign = *curr;
            curr++;
        } else if (end_not_reached && is_digit(*curr)) {
        } else {
            return false;  // empty E is not allowed
        }
        read = 0;
        end_not_reached = curr != s_end;
        while (end_not_reached && is_digit(*curr)) {
            exponent *= 10;
            exponent += (int)(*curr - '0');
            curr++;
            read++;
            end_not_reached = curr != s_end;
        }
        exponent *= exp_sign == '+' ? 1 : -1;
        if (read == 0) {
            return false;
        }
    }
    *result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    return true;
}

inline bool is_blank(char c)
{
    return c == ' ' || c == '\t';
}

// parseReal (tiny_obj_loader.h:680-689): skip blanks, the token runs to the next blank / '\r' / end of line
float parse_real(const char *&tok, const char *line_end, double default_value = 0.0)
{
    while (tok < line_end && is_blank(*tok)) {
        ++tok;
    }
    const char *end = tok;
    while (end < line_end && !is_blank(*end) && *end != '\r') {
        ++end;
    }
    double val = default_value;
    try_parse_double(tok, end, &val);
    tok = end;
    return (float)val;
}

// atoi on a bounded buffer
int parse_int(const char *tok, const char *line_end)
{
    while (tok < line_end && (is_blank(*tok) || *tok == '\n' || *tok == '\v' || *tok == '\f' || *tok == '\r')) {
        ++tok;
    }
    bool neg = false;
    if (tok < line_end && (*tok == '+' || *tok == '-')) {
        neg = *tok == '-';
        ++tok;
    }
    long long v = 0;
    while (tok < line_end && is_digit(*tok)) {
        v = v * 10 + (*tok - '0');
        if (v > 0x7fffffffll) {
            v = 0x7fffffffll;
        }
        ++tok;
    }
    return (int)(neg ? -v : v);
}

// fixIndex (tiny_obj_loader.h): 1-based -> 0-based, negative = relative to the elements read so far, 0 is invalid
inline bool fix_index(int idx, int n, int *ret)
{
    if (idx > 0) {
        *ret = idx - 1;
        return true;
    }
    if (idx == 0) {
        return false;
    }
    *ret = n + idx;
    return true;
}

inline const char *skip_to_slash_or_blank(const char *tok, const char *line_end)
{
    while (tok < line_end && *tok != '/' && !is_blank(*tok) && *tok != '\r') {
        ++tok;
    }
    return tok;
}

struct Triple {
    int v, vn, vt;
};

// parseTriple (tiny_obj_loader.h:823-877): i, i/j, i//k, i/j/k
bool parse_triple(const char *&tok, const char *line_end, int vsize, int vnsize, int vtsize, Triple *out)
{
    Triple t{-1, -1, -1};
    if (!fix_index(parse_int(tok, line_end), vsize, &t.v)) {
        return false;
    }
    tok = skip_to_slash_or_blank(tok, line_end);
    if (tok >= line_end || *tok != '/') {
        *out = t;
        return true;
    }
    ++tok;
    if (tok < line_end && *tok == '/') {  // i//k
        ++tok;
        if (!fix_index(parse_int(tok, line_end), vnsize, &t.vn)) {
            return false;
        }
        tok = skip_to_slash_or_blank(tok, line_end);
        *out = t;
        return true;
    }
    if (!fix_index(parse_int(tok, line_end), vtsize, &t.vt)) {
        return false;
    }
    tok = skip_to_slash_or_blank(tok, line_end);
    if (tok >= line_end || *tok != '/') {
        *out = t;
        return true;
    }
    ++tok;
    if (!fix_index(parse_int(tok, line_end), vnsize, &t.vn)) {
        return false;
    }
    tok = skip_to_slash_or_blank(tok, line_end);
    *out = t;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
struct MappedFile {
    const char *data = nullptr;
    size_t size = 0;
    int fd = -1;
    explicit MappedFile(const std::string &path, bool sequential = true)
    {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) {
            throw std::runtime_error("cannot open " + path);
        }
        struct stat st;
        if (fstat(fd, &st) != 0) {
            close(fd);
            throw std::runtime_error("cannot stat " + path);
        }
        size = (size_t)st.st_size;
        if (size) {
            void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) {
                close(fd);
                throw std::runtime_error("cannot map " + path);
            }
            madvise(p, size, sequential ? MADV_SEQUENTIAL : MADV_WILLNEED);
            data = static_cast<const char *>(p);
        }
    }
    ~MappedFile()
    {
        if (data) {
            munmap(const_cast<char *>(data), size);
        }
        if (fd >= 0) {
            close(fd);
        }
    }
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
};

enum EventKind { kEvGroup, kEvObject, kEvUseMtl, kEvMtlLib };
struct Event {
    EventKind kind;
    size_t face;       // triangle slots of the faces read before this line (chunk-local; + Chunk::f0 = global)
    size_t nv;         // `v` statements read before this line (chunk-local; + Chunk::v0 = global)
    std::string text;  // the rest of the line
};

// A face with more than three corners: its corners wait in Chunk::poly_corners until every vertex position is known
struct Polygon {
    size_t slot;     // first of its corners - 2 triangle slots (global)
    uint32_t first;  // index of its first corner in Chunk::poly_corners
    uint32_t corners;
};

constexpr int32_t kNoTriangle = INT32_MIN;  // v index of a triangle slot that holds no triangle

struct Chunk {
    size_t begin = 0, end = 0;
    size_t nv = 0, nvt = 0, nvn = 0, nf = 0;       // counts of this chunk (nf: triangle slots, face_slots() per face)
    size_t v0 = 0, vt0 = 0, vn0 = 0, f0 = 0;       // counts before this chunk
    std::vector<Event> events;
    std::vector<Polygon> polys;
    std::vector<Triple> poly_corners;
    bool empty_slots = false;  // a face of fewer than three corners was read (its slot stays empty)
    std::string error;
};

enum LineKind { kLineOther, kLineV, kLineVT, kLineVN, kLineF, kLineG, kLineO, kLineUseMtl, kLineMtlLib };

// the statement of a line after tinyobjloader's own trimming (LoadObj, tiny_obj_loader.h:1845-1870)
inline LineKind classify(const char *&tok, const char *line_end)
{
    while (tok < line_end && is_blank(*tok)) {
        ++tok;
    }
    const size_t n = (size_t)(line_end - tok);
    if (n < 2) {
        return kLineOther;
    }
    const char c0 = tok[0], c1 = tok[1];
    if (c0 == 'v') {
        if (is_blank(c1)) {
            tok += 2;
            return kLineV;
        }
        if (n >= 3 && is_blank(tok[2])) {
            if (c1 == 't') {
                tok += 3;
                return kLineVT;
            }
            if (c1 == 'n') {
                tok += 3;
                return kLineVN;
            }
        }
        return kLineOther;
    }
    if (c0 == 'f' && is_blank(c1)) {
        tok += 2;
        return kLineF;
    }
    if (c0 == 'g' && is_blank(c1)) {
        return kLineG;
    }
    if (c0 == 'o' && is_blank(c1)) {
        tok += 2;
        return kLineO;
    }
    if (n >= 7 && is_blank(tok[6])) {
        if (std::memcmp(tok, "usemtl", 6) == 0) {
            tok += 7;
            return kLineUseMtl;
        }
        if (std::memcmp(tok, "mtllib", 6) == 0) {
            tok += 7;
            return kLineMtlLib;
        }
    }
    return kLineOther;
}

// Corners of an `f` statement = its blank-separated tokens: LoadObj (:1950-1973) reads one index triple, then skips blanks
// and carriage returns, until the line ends; a token that is more or less than one triple makes parseTriple fail (the
// second pass reports that).
inline size_t count_corners(const char *tok, const char *line_end)
{
    size_t n = 0;
    bool in_token = false;
    for (; tok < line_end; ++tok) {
        const bool blank = is_blank(*tok) || *tok == '\r';
        n += (!blank && !in_token) ? 1 : 0;
        in_token = !blank;
    }
    return n;
}

// Triangle slots a face of `corners` corners gets: the ear clipping of exportGroupsToShape emits at most corners - 2
// triangles; a face of fewer than three corners still is a face of its group (faceGroup.empty() decides whether a shape
// is exported), so it keeps one slot, which stays empty.
inline size_t face_slots(size_t corners)
{
    return corners > 3 ? corners - 2 : 1;
}

// [line_begin, line_end) of the line that starts at p; line_end excludes '\n' and a '\r' before it
inline const char *line_end_of(const char *p, const char *chunk_end, const char *&next)
{
    const char *nl = static_cast<const char *>(std::memchr(p, '\n', (size_t)(chunk_end - p)));
    const char *e = nl ? nl : chunk_end;
    next = nl ? nl + 1 : chunk_end;
    if (e > p && e[-1] == '\r') {
        --e;
    }
    return e;
}

struct Material {  // the fields of tinyobj::material_t that Scene::load_obj reads
    std::string name;
    float diffuse[3] = {0.f, 0.f, 0.f};  // InitMaterial, tiny_obj_loader.h:1039
    float shininess = 1.f;               // :1046
    std::string diffuse_texname;
};

// LoadMtl (tiny_obj_loader.h:1353-1723), the statements Scene::load_obj depends on
void load_mtl(const std::string &path, std::vector<Material> &materials, std::map<std::string, int> &material_map, bool &found)
{
    std::ifstream in(path.c_str());
    found = (bool)in;
    if (!found) {
        return;
    }
    Material material;
    std::string linebuf;
    while (std::getline(in, linebuf)) {
        if (!linebuf.empty()) {
            linebuf = linebuf.substr(0, linebuf.find_last_not_of(" \t") + 1);  // trailing blanks (:1374)
        }
        if (!linebuf.empty() && linebuf.back() == '\n') {
            linebuf.pop_back();
        }
        if (!linebuf.empty() && linebuf.back() == '\r') {
            linebuf.pop_back();
        }
        if (linebuf.empty()) {
            continue;
        }
        const char *tok = linebuf.c_str();
        const char *end = tok + linebuf.size();
        while (tok < end && is_blank(*tok)) {
            ++tok;
        }
        if (tok >= end || *tok == '#') {
            continue;
        }
        const size_t n = (size_t)(end - tok);
        if (n >= 7 && std::memcmp(tok, "newmtl", 6) == 0 && is_blank(tok[6])) {
            if (!material.name.empty()) {  // flush the previous material (:1407-1411)
                material_map.insert(std::make_pair(material.name, (int)materials.size()));
                materials.push_back(material);
            }
            material = Material();
            material.name = std::string(tok + 7, end);
            continue;
        }
        if (n >= 3 && tok[0] == 'K' && tok[1] == 'd' && is_blank(tok[2])) {
            tok += 2;
            material.diffuse[0] = parse_real(tok, end);
            material.diffuse[1] = parse_real(tok, end);
            material.diffuse[2] = parse_real(tok, end);
            continue;
        }
        if (n >= 3 && tok[0] == 'N' && tok[1] == 's' && is_blank(tok[2])) {
            tok += 2;
            material.shininess = parse_real(tok, end);
            continue;
        }
        if (n >= 7 && std::memcmp(tok, "map_Kd", 6) == 0 && is_blank(tok[6])) {
            // ParseTextureNameAndOption (:906-985): options with their arguments, then the name = the rest of the line (it may
            // hold blanks)
            tok += 7;
            std::string name;
            static const struct {
                const char *flag;
                int args;
            } options[] = {{"-blendu", 1}, {"-blendv", 1}, {"-clamp", 1}, {"-boost", 1}, {"-bm", 1}, {"-o", 3}, {"-s", 3}, {"-t", 3},
                           {"-type", 1}, {"-imfchan", 1}, {"-mm", 2}, {"-colorspace", 1}};
            while (tok < end) {
                while (tok < end && is_blank(*tok)) {
                    ++tok;
                }
                bool is_option = false;
                for (const auto &o : options) {
                    const size_t len = std::strlen(o.flag);
                    if ((size_t)(end - tok) > len && std::memcmp(tok, o.flag, len) == 0 && is_blank(tok[len])) {
                        tok += len;
                        for (int a = 0; a < o.args; ++a) {  // each argument: blanks, then a run of non-blanks (none at the line end)
                            while (tok < end && is_blank(*tok)) {
                                ++tok;
                            }
                            while (tok < end && !is_blank(*tok)) {
                                ++tok;
                            }
                        }
                        is_option = true;
                        break;
                    }
                }
                if (!is_option) {
                    name = std::string(tok, end);
                    break;
                }
            }
            material.diffuse_texname = name;
            continue;
        }
    }
    // the last material is flushed whatever its name (:1716-1719)
    material_map.insert(std::make_pair(material.name, (int)materials.size()));
    materials.push_back(material);
}

using crt_image::decode_image_rgba;

void load_image_rgba_flipped(const std::string &path, std::vector<uint8_t> &out, int &width, int &height)
{
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in) {
        throw std::runtime_error("Failed to load " + path);  // util/material.cpp:11-13
    }
    const std::vector<uint8_t> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    decode_image_rgba(file.data(), file.size(), path, out, width, height, true);
}

// glm::normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x)
void normalize3(float v[3])
{
    const float inv = 1.f / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= inv, v[1] *= inv, v[2] *= inv;
}
void cross3(const float a[3], const float b[3], float out[3])
{
    out[0] = a[1] * b[2] - b[1] * a[2];
    out[1] = a[2] * b[0] - b[2] * a[0];
    out[2] = a[0] * b[1] - b[0] * a[1];
}
// ortho_basis (util/util.cpp: the host twin of backends/embree/util.ih:24-46)
void ortho_basis(float v_x[3], float v_y[3], const float n[3])
{
    v_y[0] = v_y[1] = v_y[2] = 0.f;
    if (n[0] < 0.6f && n[0] > -0.6f) {
        v_y[0] = 1.f;
    } else if (n[1] < 0.6f && n[1] > -0.6f) {
        v_y[1] = 1.f;
    } else if (n[2] < 0.6f && n[2] > -0.6f) {
        v_y[2] = 1.f;
    } else {
        v_y[0] = 1.f;
    }
    cross3(v_y, n, v_x);
    normalize3(v_x);
    cross3(n, v_x, v_y);
    normalize3(v_y);
}

struct GeometryData {
    std::vector<float> vertices, uvs;
    std::vector<uint32_t> indices;
};

}  // namespace

struct crtio_scene {
    bool white_diffuse = false;  // MaterialMode::WHITE_DIFFUSE (util/scene.h:21): no materials are read, every geometry gets the default one
    std::unique_ptr<MappedFile> mapping;  // .crts: the geometry arrays are the file's own bytes where their alignment allows
    std::vector<GeometryData> geometries;  // OBJ: the arrays of its shapes
    std::vector<std::vector<uint8_t>> unaligned_copies;  // .crts / glTF: arrays that cannot be used where they are in the file
    std::vector<std::unique_ptr<MappedFile>> buffer_mappings;  // glTF: external .bin files
    std::vector<std::vector<uint8_t>> decoded_buffers;         // glTF: buffers given as data: URIs
    std::vector<crt_geometry_t> geometry_views;
    std::vector<crt_mesh_t> meshes;
    std::vector<std::vector<uint32_t>> material_ids;  // per parameterized mesh
    std::vector<crt_parameterized_mesh_t> parameterized_meshes;
    std::vector<crt_instance_t> instances;
    std::vector<crt_material_t> materials;
    std::vector<std::vector<uint8_t>> texture_data;
    std::vector<crt_image_t> textures;
    std::vector<std::string> texture_names;  // Image::name
    std::vector<crt_quad_light_t> lights;
    std::vector<crtio_camera_t> cameras;
    crt_scene_t view{};
    std::string warnings;
    double timings[4] = {0, 0, 0, 0};

    void finish_views()  // after every vector above has its final size
    {
        for (size_t i = 0; i < parameterized_meshes.size(); ++i) {
            parameterized_meshes[i].material_ids = material_ids[i].data();
            parameterized_meshes[i].num_material_ids = (uint32_t)material_ids[i].size();
        }
        view = crt_scene_t{meshes.data(), parameterized_meshes.data(), instances.data(), materials.data(), textures.data(), lights.data(),
                           (uint32_t)meshes.size(), (uint32_t)parameterized_meshes.size(), (uint32_t)instances.size(),
                           (uint32_t)materials.size(), (uint32_t)textures.size(), (uint32_t)lights.size(), 1u};
    }
};

namespace {

// Scene::validate_materials (scene.cpp:935-957): geometries without a material get a default DisneyMaterial
void validate_materials(crtio_scene &S, std::ostream &warn)
{
    bool need_default = false;
    for (const std::vector<uint32_t> &ids : S.material_ids) {
        need_default = need_default || std::find(ids.begin(), ids.end(), 0xffffffffu) != ids.end();
    }
    if (!need_default) {
        return;
    }
    crt_material_t d;  // DisneyMaterial's defaults, util/material.h:29-46
    std::memset(&d, 0, sizeof(d));
    d.base_color[0] = d.base_color[1] = d.base_color[2] = 0.9f;
    d.roughness = 1.f;
    d.ior = 1.5f;
    const uint32_t id = (uint32_t)S.materials.size();
    S.materials.push_back(d);
    for (std::vector<uint32_t> &ids : S.material_ids) {
        std::replace(ids.begin(), ids.end(), 0xffffffffu, id);
    }
    warn << "No materials assigned for some objects, generating a default\n";
}

// The light the loaders generate for a scene that has none (scene.cpp:216-227 emits 20, :612-624 emits 10)
crt_quad_light_t generated_light(float emission)
{
    float n[3] = {0.5f, -0.8f, -0.5f};
    normalize3(n);
    crt_quad_light_t L;
    std::memset(&L, 0, sizeof(L));
    L.emission[0] = L.emission[1] = L.emission[2] = L.emission[3] = emission;
    L.normal[0] = n[0], L.normal[1] = n[1], L.normal[2] = n[2], L.normal[3] = 0.f;
    for (int k = 0; k < 4; ++k) {
        L.position[k] = -10.f * L.normal[k];
    }
    ortho_basis(L.v_x, L.v_y, n);
    L.width = 5.f;
    L.height = 5.f;
    return L;
}

// One geometry of Scene::load_obj (util/scene.cpp:116-181): index triples -> single indices in order of first use
void remap_shape(const std::vector<float> &V, const std::vector<float> &VT, const int32_t *faces, size_t num_faces, GeometryData &g,
                 const std::string &what)
{
    size_t cap = 64;
    while (cap < num_faces * 3 * 2) {
        cap <<= 1;
    }
    struct Slot {
        int32_t v, vn, vt;
        uint32_t index;
    };
    std::vector<Slot> table(cap, Slot{-2, 0, 0, 0});
    g.indices.reserve(num_faces * 3);
    g.vertices.reserve(num_faces * 3);
    const size_t nv = V.size() / 3, nvt = VT.size() / 2;
    bool any_uv = false, any_without_uv = false;
    for (size_t c = 0; c < num_faces * 3; ++c) {
        if (c % 3 == 0 && faces[3 * c] == kNoTriangle) {  // a slot no triangle was made for
            c += 2;
            continue;
        }
        const int32_t v = faces[3 * c], vn = faces[3 * c + 1], vt = faces[3 * c + 2];
        uint64_t h = (uint64_t)(uint32_t)v * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)vt + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= ((uint64_t)(uint32_t)vn + 0x165667B1ull) * 0x27D4EB2F165667C5ull;
        size_t slot = (size_t)(h ^ (h >> 29)) & (cap - 1);
        for (;;) {
            Slot &s = table[slot];
            if (s.v == -2) {
                if (v < 0 || (size_t)v >= nv) {
                    throw std::runtime_error("vertex index out of range in " + what);
                }
                s = Slot{v, vn, vt, (uint32_t)(g.vertices.size() / 3)};
                g.vertices.push_back(V[3 * (size_t)v]);
                g.vertices.push_back(V[3 * (size_t)v + 1]);
                g.vertices.push_back(V[3 * (size_t)v + 2]);
                if (vt != -1) {
                    if (vt < 0 || (size_t)vt >= nvt) {
                        throw std::runtime_error("texture coordinate index out of range in " + what);
                    }
                    g.uvs.push_back(VT[2 * (size_t)vt]);
                    g.uvs.push_back(VT[2 * (size_t)vt + 1]);
                    any_uv = true;
                } else {
                    any_without_uv = true;
                }
                g.indices.push_back(s.index);
                break;
            }
            if (s.v == v && s.vn == vn && s.vt == vt) {
                g.indices.push_back(s.index);
                break;
            }
            slot = (slot + 1) & (cap - 1);
        }
    }
    if (any_uv && any_without_uv) {
        throw std::runtime_error("some corners of " + what + " have texture coordinates and some have none");
    }
}

// pnpoly as tinyobjloader uses it on one candidate ear (tiny_obj_loader.h:1064-1076): is (tx, ty) inside the triangle?
inline bool point_in_triangle(const float x[3], const float y[3], float tx, float ty)
{
    bool inside = false;
    for (int i = 0, j = 2; i < 3; j = i++) {
        if (((y[i] > ty) != (y[j] > ty)) && (tx < (x[j] - x[i]) * (ty - y[i]) / (y[j] - y[i]) + x[i])) {
            inside = !inside;
        }
    }
    return inside;
}

// The triangulation of one face of more than three corners, as exportGroupsToShape does it (tiny_obj_loader.h:1107-1310;
// Scene::load_obj calls LoadObj with triangulate = true): the polygon is projected on the two axes its first non-degenerate
// corner spans best, its signed area gives the winding, and ears are clipped — a corner that turns the polygon's way and
// holds no other remaining corner inside — starting the search where the last ear was found. A polygon the search gets
// stuck on (no ear in a full round) keeps the triangles found so far and loses the rest, as in tinyobjloader. All in
// float, in tinyobjloader's order of operations, since which ear comes first decides the triangles and their order.
// V / vcount: the positions read when the face group was exported (corners that point past them count as (0, 0) or are
// skipped, as there). Writes up to n - 2 triangles of 3 x (v, vn, vt) to `out`, returns how many.
size_t ear_clip(const float *V, size_t vcount, const Triple *corner, size_t n, int32_t *out)
{
    const auto known = [&](int v) { return (size_t)(int64_t)v < vcount; };
    size_t axes[2] = {1, 2};
    for (size_t k = 0; k < n; ++k) {
        const int a = corner[k % n].v, b = corner[(k + 1) % n].v, c = corner[(k + 2) % n].v;
        if (!known(a) || !known(b) || !known(c)) {
            continue;
        }
        const float *p0 = V + 3 * (size_t)a, *p1 = V + 3 * (size_t)b, *p2 = V + 3 * (size_t)c;
        const float e0x = p1[0] - p0[0], e0y = p1[1] - p0[1], e0z = p1[2] - p0[2];
        const float e1x = p2[0] - p1[0], e1y = p2[1] - p1[1], e1z = p2[2] - p1[2];
        const float cx = std::fabs(e0y * e1z - e0z * e1y);
        const float cy = std::fabs(e0z * e1x - e0x * e1z);
        const float cz = std::fabs(e0x * e1y - e0y * e1x);
        const float epsilon = std::numeric_limits<float>::epsilon();
        if (cx > epsilon || cy > epsilon || cz > epsilon) {
            if (!(cx > cy && cx > cz)) {
                axes[0] = 0;
                if (cz > cx && cz > cy) {
                    axes[1] = 1;
                }
            }
            break;
        }
    }
    float area = 0.f;
    for (size_t k = 0; k < n; ++k) {
        const int a = corner[k].v, b = corner[(k + 1) % n].v;
        if (!known(a) || !known(b)) {
            continue;
        }
        const float ax = V[3 * (size_t)a + axes[0]], ay = V[3 * (size_t)a + axes[1]];
        const float bx = V[3 * (size_t)b + axes[0]], by = V[3 * (size_t)b + axes[1]];
        area += (ax * by - ay * bx) * 0.5f;
    }
    const auto emit = [&](const Triple &a, const Triple &b, const Triple &c, size_t at) {
        const Triple t[3] = {a, b, c};
        for (int k = 0; k < 3; ++k) {
            out[9 * at + 3 * k] = t[k].v;
            out[9 * at + 3 * k + 1] = t[k].vn;
            out[9 * at + 3 * k + 2] = t[k].vt;
        }
    };
    Triple small[16];
    std::vector<Triple> large;
    Triple *rest = small;  // the corners not clipped yet
    if (n > 16) {
        large.assign(corner, corner + n);
        rest = large.data();
    } else {
        std::copy(corner, corner + n, small);
    }
    size_t left = n, emitted = 0, guess = 0;
    size_t rounds_left = n, left_before = n;  // iterations allowed without clipping a corner
    while (left > 3 && rounds_left > 0) {
        if (guess >= left) {
            guess -= left;
        }
        if (left_before != left) {
            left_before = left;
            rounds_left = left;
        } else {
            --rounds_left;
        }
        Triple ind[3];
        float x[3], y[3];
        for (size_t k = 0; k < 3; ++k) {
            ind[k] = rest[(guess + k) % left];
            if (known(ind[k].v)) {
                x[k] = V[3 * (size_t)ind[k].v + axes[0]];
                y[k] = V[3 * (size_t)ind[k].v + axes[1]];
            } else {
                x[k] = 0.f;
                y[k] = 0.f;
            }
        }
        const float e0x = x[1] - x[0], e0y = y[1] - y[0], e1x = x[2] - x[1], e1y = y[2] - y[1];
        const float cross = e0x * e1y - e0y * e1x;
        if (cross * area < 0.f) {  // turns against the polygon: not an ear
            ++guess;
            continue;
        }
        bool overlap = false;
        for (size_t other = 3; other < left && !overlap; ++other) {
            const int ov = rest[(guess + other) % left].v;
            if (known(ov)) {
                overlap = point_in_triangle(x, y, V[3 * (size_t)ov + axes[0]], V[3 * (size_t)ov + axes[1]]);
            }
        }
        if (overlap) {
            ++guess;
            continue;
        }
        emit(ind[0], ind[1], ind[2], emitted++);
        for (size_t r = (guess + 1) % left; r + 1 < left; ++r) {  // the ear's tip leaves the polygon
            rest[r] = rest[r + 1];
        }
        --left;
    }
    if (left == 3) {
        emit(rest[0], rest[1], rest[2], emitted++);
    }
    return emitted;
}

void load_obj_impl(const std::string &file, int threads, crtio_scene &S)
{
    const double t_start = now_s();
    const unsigned nthreads = crt::host_threads(threads);
    MappedFile map(file);
    // ---- chunks at line boundaries ----
    const size_t target = std::max<size_t>((size_t)1 << 20, map.size / (std::max(1u, nthreads) * 8u) + 1);
    std::vector<Chunk> chunks;
    for (size_t b = 0; b < map.size;) {
        size_t e = std::min(map.size, b + target);
        if (e < map.size) {
            const void *nl = std::memchr(map.data + e, '\n', map.size - e);
            e = nl ? (size_t)(static_cast<const char *>(nl) - map.data) + 1 : map.size;
        }
        Chunk c;
        c.begin = b;
        c.end = e;
        chunks.push_back(c);
        b = e;
    }
    // ---- pass 1: count statements, collect the shape / material events ----
    crt::parallel_blocks((uint32_t)chunks.size(), nthreads, [&](uint32_t ci) {
        Chunk &c = chunks[ci];
        const char *p = map.data + c.begin, *chunk_end = map.data + c.end;
        while (p < chunk_end) {
            const char *next;
            const char *le = line_end_of(p, chunk_end, next);
            const char *tok = p;
            switch (classify(tok, le)) {
            case kLineV: c.nv++; break;
            case kLineVT: c.nvt++; break;
            case kLineVN: c.nvn++; break;
            case kLineF: c.nf += face_slots(count_corners(tok, le)); break;
            case kLineG: c.events.push_back(Event{kEvGroup, c.nf, c.nv, std::string()}); break;
            case kLineO: c.events.push_back(Event{kEvObject, c.nf, c.nv, std::string()}); break;
            case kLineUseMtl: c.events.push_back(Event{kEvUseMtl, c.nf, c.nv, std::string(tok, le)}); break;
            case kLineMtlLib: c.events.push_back(Event{kEvMtlLib, c.nf, c.nv, std::string(tok, le)}); break;
            default: break;
            }
            p = next;
        }
    });
    size_t nv = 0, nvt = 0, nvn = 0, nf = 0;
    for (Chunk &c : chunks) {
        c.v0 = nv, c.vt0 = nvt, c.vn0 = nvn, c.f0 = nf;
        nv += c.nv, nvt += c.nvt, nvn += c.nvn, nf += c.nf;
    }
    if (nv >= 0x7fffffffull || nvt >= 0x7fffffffull || nvn >= 0x7fffffffull || nf >= 0x7fffffffull / 3) {
        throw std::runtime_error("OBJ file too large for 32-bit indices: " + file);
    }
    // ---- pass 2: parse into the global arrays ----
    std::vector<float> V(nv * 3), VT(nvt * 2);
    std::vector<int32_t> F(nf * 9);  // per face 3 x (v, vn, vt)
    crt::parallel_blocks((uint32_t)chunks.size(), nthreads, [&](uint32_t ci) {
        Chunk &c = chunks[ci];
        size_t v = c.v0, vt = c.vt0, vn = c.vn0, f = c.f0;
        const char *p = map.data + c.begin, *chunk_end = map.data + c.end;
        while (p < chunk_end && c.error.empty()) {
            const char *next;
            const char *le = line_end_of(p, chunk_end, next);
            const char *tok = p;
            switch (classify(tok, le)) {
            case kLineV:  // parseVertexWithColor (:733-754): x y z, then optional colours that do not matter here
                V[3 * v] = parse_real(tok, le);
                V[3 * v + 1] = parse_real(tok, le);
                V[3 * v + 2] = parse_real(tok, le);
                ++v;
                break;
            case kLineVT:  // parseReal2 (:1900-1907)
                VT[2 * vt] = parse_real(tok, le);
                VT[2 * vt + 1] = parse_real(tok, le);
                ++vt;
                break;
            case kLineVN:
                ++vn;
                break;
            case kLineF: {
                const size_t expected = count_corners(tok, le);
                while (tok < le && is_blank(*tok)) {
                    ++tok;
                }
                size_t corners = 0;
                Triple t[3];
                const size_t poly_first = c.poly_corners.size();
                while (tok < le && *tok != '\r') {
                    Triple tr;
                    if (!parse_triple(tok, le, (int)v, (int)vn, (int)vt, &tr)) {
                        c.error = "Failed parse `f' line(e.g. zero value for face index)";  // :1945-1952
                        break;
                    }
                    if (corners < 3) {
                        t[corners] = tr;
                    }
                    if (expected > 3) {
                        c.poly_corners.push_back(tr);
                    }
                    ++corners;
                    while (tok < le && (is_blank(*tok) || *tok == '\r')) {
                        ++tok;
                    }
                }
                if (c.error.empty() && corners != expected) {
                    c.error = "Failed parse `f' line(a corner that is not one index triple)";
                }
                if (c.error.empty()) {
                    const size_t slots = face_slots(corners);
                    for (size_t k = 0; k < slots; ++k) {
                        F[9 * (f + k)] = kNoTriangle;
                    }
                    if (corners == 3) {
                        for (int k = 0; k < 3; ++k) {
                            F[9 * f + 3 * k] = t[k].v;
                            F[9 * f + 3 * k + 1] = t[k].vn;
                            F[9 * f + 3 * k + 2] = t[k].vt;
                        }
                    } else if (corners > 3) {
                        c.polys.push_back(Polygon{f, (uint32_t)poly_first, (uint32_t)corners});
                    } else {
                        c.empty_slots = true;
                    }
                    f += slots;
                }
                break;
            }
            default: break;
            }
            p = next;
        }
    });
    for (const Chunk &c : chunks) {
        if (!c.error.empty()) {
            throw std::runtime_error("TinyOBJ Error loading " + file + " error: " + c.error);
        }
    }
    // ---- shapes, from the events in file order (LoadObj's handling of usemtl / g / o, :1979-2128 and :2200-2212) ----
    struct Segment {   // one exportGroupsToShape call
        size_t begin, end;  // triangle slots
        int material;
        size_t vcount;      // positions read when the group was exported (what the ear clipping may look at)
    };
    struct Shape {
        std::vector<Segment> segments;
    };
    std::vector<Shape> shapes;
    std::vector<Material> obj_materials;
    std::map<std::string, int> material_map;
    std::string obj_base_dir = file.substr(0, file.rfind('/'));  // util/scene.cpp:105 (the whole name if there is no '/')
    std::string mtl_base = obj_base_dir;
    if (!mtl_base.empty() && mtl_base.back() != '/') {
        mtl_base += "/";
    }
    std::ostringstream warn;
    {
        Shape shape;
        int material = -1;
        size_t group_begin = 0;  // first face of the current face group
        auto export_group = [&](size_t pos, size_t vcount) {
            if (group_begin == pos) {
                return false;
            }
            shape.segments.push_back(Segment{group_begin, pos, material, vcount});
            group_begin = pos;
            return true;
        };
        for (const Chunk &c : chunks) {
            for (const Event &ev : c.events) {
                const size_t pos = c.f0 + ev.face, vcount = c.v0 + ev.nv;
                switch (ev.kind) {
                case kEvUseMtl: {
                    const auto it = material_map.find(ev.text);
                    const int id = it != material_map.end() ? it->second : -1;
                    if (id != material) {
                        export_group(pos, vcount);
                        material = id;
                    }
                    break;
                }
                case kEvMtlLib: {
                    std::vector<std::string> names;
                    std::stringstream ss(ev.text);
                    std::string item;
                    while (std::getline(ss, item, ' ')) {  // SplitString(token, ' ')
                        names.push_back(item);
                    }
                    bool found = false;
                    for (const std::string &n : names) {
                        load_mtl(mtl_base + n, obj_materials, material_map, found);
                        if (found) {
                            break;
                        }
                        warn << "Material file [ " << mtl_base + n << " ] not found.\n";
                    }
                    if (!found) {
                        warn << "Failed to load material file(s). Use default material.\n";
                    }
                    break;
                }
                case kEvGroup:
                    export_group(pos, vcount);
                    if (!shape.segments.empty()) {
                        shapes.push_back(shape);
                    }
                    shape = Shape();
                    break;
                case kEvObject:
                    if (export_group(pos, vcount)) {  // (a shape whose faces were all flushed by an earlier usemtl is dropped: :2105-2109)
                        shapes.push_back(shape);
                    }
                    shape = Shape();
                    break;
                }
            }
        }
        const bool ret = export_group(nf, nv);
        if (ret || !shape.segments.empty()) {
            shapes.push_back(shape);
        }
    }
    // ---- faces of more than three corners: ear clipping into their triangle slots, now that the positions are there ----
    bool empty_slots = false;
    {
        std::vector<const Segment *> exported;  // in slot order
        size_t num_polys = 0;
        for (const Shape &sh : shapes) {
            for (const Segment &seg : sh.segments) {
                exported.push_back(&seg);
            }
        }
        for (const Chunk &c : chunks) {
            num_polys += c.polys.size();
            empty_slots = empty_slots || c.empty_slots;
        }
        if (num_polys) {
            std::vector<uint8_t> chunk_lost(chunks.size(), 0);
            crt::parallel_blocks((uint32_t)chunks.size(), nthreads, [&](uint32_t ci) {
                const Chunk &c = chunks[ci];
                for (const Polygon &poly : c.polys) {
                    // the face group this polygon was exported with (none: its shape was dropped)
                    auto it = std::upper_bound(exported.begin(), exported.end(), poly.slot,
                                               [](size_t slot, const Segment *seg) { return slot < seg->end; });
                    if (it == exported.end() || poly.slot < (*it)->begin) {
                        continue;
                    }
                    const size_t made = ear_clip(V.data(), (*it)->vcount, c.poly_corners.data() + poly.first, poly.corners, F.data() + 9 * poly.slot);
                    if (made != (size_t)poly.corners - 2) {
                        chunk_lost[ci] = 1;
                    }
                }
            });
            empty_slots = empty_slots || std::find(chunk_lost.begin(), chunk_lost.end(), 1) != chunk_lost.end();
        }
    }
    const double t_parsed = now_s();
    // ---- Scene::load_obj: one geometry per shape ----
    S.geometries.resize(shapes.size());
    S.material_ids.assign(1, std::vector<uint32_t>(shapes.size()));
    std::vector<uint32_t> &shape_material = S.material_ids[0];
    for (size_t s = 0; s < shapes.size(); ++s) {
        // The material of a shape is its first triangle's (scene.cpp:127); a shape whose triangles do not all have that
        // material gets the warning of scene.cpp:131-137. Only face groups that hold a triangle count.
        bool first = true, mixed = false;
        int material = -1;
        for (const Segment &seg : shapes[s].segments) {
            bool holds_triangle = !empty_slots;
            for (size_t f = seg.begin; f < seg.end && !holds_triangle; ++f) {
                holds_triangle = F[9 * f] != kNoTriangle;
            }
            if (!holds_triangle) {
                continue;
            }
            if (first) {
                material = seg.material;
                first = false;
            } else if (seg.material != material) {
                mixed = true;
            }
        }
        if (first) {  // (the reference reads material_ids[0] of an empty array here)
            throw std::runtime_error("a shape without a triangle (its faces have fewer than three corners or no area) in " + file);
        }
        shape_material[s] = S.white_diffuse ? 0xffffffffu : (uint32_t)material;  // scene.cpp:126-130
        if (mixed) {
            warn << "Warning: per-face material IDs are not supported, materials may look wrong. Please reexport your mesh with each "
                    "material group as an OBJ group\n";
        }
    }
    std::vector<std::string> shape_errors(shapes.size());
    crt::parallel_blocks((uint32_t)shapes.size(), nthreads, [&](uint32_t s) {
        try {
            const size_t begin = shapes[s].segments.front().begin, end = shapes[s].segments.back().end;
            remap_shape(V, VT, F.data() + 9 * begin, end - begin, S.geometries[s], file + " shape " + std::to_string(s));
        } catch (const std::exception &e) {
            shape_errors[s] = e.what();
        }
    });
    for (const std::string &e : shape_errors) {
        if (!e.empty()) {
            throw std::runtime_error(e);
        }
    }
    const double t_remapped = now_s();
    // ---- materials (scene.cpp:188-214) ----
    std::map<std::string, int32_t> texture_ids;
    std::vector<std::string> texture_files;
    if (S.white_diffuse) {
        obj_materials.clear();  // scene.cpp:188
    }
    for (const Material &m : obj_materials) {
        crt_material_t d;
        std::memset(&d, 0, sizeof(d));
        d.base_color[0] = m.diffuse[0], d.base_color[1] = m.diffuse[1], d.base_color[2] = m.diffuse[2];
        d.ior = 1.5f;  // DisneyMaterial's defaults, util/material.h:29-46
        const float spec = m.shininess / 500.f;
        d.specular = spec < 0.f ? 0.f : (spec > 1.f ? 1.f : spec);
        const float rough = 1.f - d.specular;
        d.roughness = rough < 0.f ? 0.f : (rough > 1.f ? 1.f : rough);
        d.specular_transmission = 0.f;
        if (!m.diffuse_texname.empty()) {
            std::string path = m.diffuse_texname;
            std::replace(path.begin(), path.end(), '\\', '/');  // canonicalize_path, util/util.cpp
            auto it = texture_ids.find(m.diffuse_texname);
            if (it == texture_ids.end()) {
                it = texture_ids.insert(std::make_pair(m.diffuse_texname, (int32_t)texture_files.size())).first;
                texture_files.push_back(obj_base_dir + "/" + path);
                S.texture_names.push_back(m.diffuse_texname);
            }
            const uint32_t tex_mask = 0x80000000u | ((uint32_t)it->second & 0x1fffffffu);  // TEXTURED_PARAM_MASK, SET_TEXTURE_ID
            std::memcpy(&d.base_color[0], &tex_mask, 4);
        }
        S.materials.push_back(d);
    }
    validate_materials(S, warn);
    S.texture_data.resize(texture_files.size());
    S.textures.resize(texture_files.size());
    std::vector<std::string> tex_errors(texture_files.size());
    crt::parallel_blocks((uint32_t)texture_files.size(), nthreads, [&](uint32_t i) {
        try {
            int w = 0, h = 0;
            load_image_rgba_flipped(texture_files[i], S.texture_data[i], w, h);
            S.textures[i] = crt_image_t{S.texture_data[i].data(), w, h, 4, CRT_COLOR_SPACE_SRGB};
        } catch (const std::exception &e) {
            tex_errors[i] = e.what();
        }
    });
    for (const std::string &e : tex_errors) {
        if (!e.empty()) {
            throw std::runtime_error(e);
        }
    }
    S.lights.push_back(generated_light(20.f));  // scene.cpp:216-227
    // ---- views: one mesh of all the geometries, one instance of it ----
    S.geometry_views.resize(S.geometries.size());
    for (size_t g = 0; g < S.geometries.size(); ++g) {
        const GeometryData &gd = S.geometries[g];
        S.geometry_views[g] = crt_geometry_t{gd.vertices.data(), gd.uvs.empty() ? nullptr : gd.uvs.data(), gd.indices.data(),
                                             (uint32_t)(gd.vertices.size() / 3), (uint32_t)(gd.indices.size() / 3)};
    }
    S.meshes.push_back(crt_mesh_t{S.geometry_views.data(), (uint32_t)S.geometry_views.size()});
    S.parameterized_meshes.push_back(crt_parameterized_mesh_t{nullptr, 0u, 0u});
    crt_instance_t instance;
    std::memset(&instance, 0, sizeof(instance));
    instance.transform[0] = instance.transform[5] = instance.transform[10] = instance.transform[15] = 1.f;
    instance.parameterized_mesh_id = 0;
    S.instances.push_back(instance);
    S.finish_views();
    S.warnings = warn.str();
    const double t_end = now_s();
    S.timings[0] = t_end - t_start;
    S.timings[1] = t_parsed - t_start;
    S.timings[2] = t_remapped - t_parsed;
    S.timings[3] = t_end - t_remapped;
}

// ---------------------------------------------------------------------------------------------------------------
// .crts (util/scene.cpp:417-625); its header is JSON (json_reader.h)
using crt_json::Json;
using crt_json::JsonParser;

// dtype_stride(parse_dtype(name)) (util/gltf_types.cpp:144-215, :431-505): "<SHAPE>_<COMPONENT>" or a scalar's long name
size_t crts_dtype_stride(const std::string &name)
{
    static const struct {
        const char *name;
        size_t bytes;
    } scalars[] = {{"INT_8", 1}, {"UINT_8", 1}, {"INT_16", 2}, {"UINT_16", 2}, {"INT_32", 4}, {"UINT_32", 4}, {"FLOAT_32", 4}, {"FLOAT_64", 8}};
    for (const auto &sc : scalars) {
        if (name == sc.name) {
            return sc.bytes;
        }
    }
    static const struct {
        const char *prefix;
        size_t components;
    } shapes[] = {{"VEC2_", 2}, {"VEC3_", 3}, {"VEC4_", 4}, {"MAT2_", 4}, {"MAT3_", 9}, {"MAT4_", 16}};
    static const struct {
        const char *suffix;
        size_t bytes;
    } components[] = {{"I8", 1}, {"U8", 1}, {"I16", 2}, {"U16", 2}, {"I32", 4}, {"U32", 4}, {"F32", 4}, {"F64", 8}};
    for (const auto &sh : shapes) {
        if (name.compare(0, 5, sh.prefix) == 0) {
            for (const auto &co : components) {
                if (name.compare(5, std::string::npos, co.suffix) == 0) {
                    return sh.components * co.bytes;
                }
            }
        }
    }
    throw std::runtime_error("Invalid data type string " + name);
}

// glm::normalize of a vec4: v * (1 / sqrt(dot)), dot = (x*x + y*y) + (z*z + w*w) (glm's compute_dot<vec<4>>)
void normalize4(const float v[4], float out[4])
{
    const float inv = 1.f / std::sqrt((v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]));
    for (int k = 0; k < 4; ++k) {
        out[k] = v[k] * inv;
    }
}

void load_crts_impl(const std::string &file, int threads, crtio_scene &S)
{
    const double t_start = now_s();
    const unsigned nthreads = crt::host_threads(threads);
    S.mapping.reset(new MappedFile(file, /*sequential=*/false));
    const MappedFile &map = *S.mapping;
    if (map.size < sizeof(uint64_t)) {
        throw std::runtime_error("not a crts file (too short): " + file);
    }
    uint64_t json_size = 0;
    std::memcpy(&json_size, map.data, sizeof(json_size));
    if (json_size > map.size - sizeof(uint64_t)) {
        throw std::runtime_error("not a crts file (header size past the end): " + file);
    }
    const Json header = JsonParser(map.data + sizeof(uint64_t), map.data + sizeof(uint64_t) + json_size).parse_document();
    if (header.kind != Json::kObject) {
        throw std::runtime_error("scene header: not an object");
    }
    const uint8_t *data_base = reinterpret_cast<const uint8_t *>(map.data) + sizeof(uint64_t) + json_size;
    const size_t data_size = map.size - sizeof(uint64_t) - (size_t)json_size;
    static const Json null_json;
    const auto section = [&](const char *key) -> const Json & {
        const Json *j = header.find(key);
        return j ? *j : null_json;
    };
    const Json &views = section("buffer_views");
    struct View {
        const uint8_t *data;
        size_t bytes;  // whole elements of the view's own type (Accessor: count = length / stride)
    };
    const auto view_of = [&](const Json &id, const std::string &where) {
        const uint64_t view_id = id.number<uint64_t>(where);
        const std::string vw = "buffer view " + std::to_string(view_id);
        const Json &v = views.at((size_t)view_id, "buffer_views");
        const size_t stride = crts_dtype_stride(v.at("type", vw).string(vw + " type"));
        const uint64_t offset = v.at("byte_offset", vw).number<uint64_t>(vw), length = v.at("byte_length", vw).number<uint64_t>(vw);
        if (offset > data_size || length > data_size - offset) {
            throw std::runtime_error("crts: " + vw + " reaches past the end of " + file);
        }
        return View{data_base + offset, (size_t)(length / stride) * stride};
    };
    // ---- meshes: one geometry each; the arrays stay where they are in the mapped file ----
    const Json &meshes = section("meshes");
    const size_t num_meshes = meshes.size();
    S.geometries.resize(num_meshes);
    S.geometry_views.resize(num_meshes);
    S.meshes.resize(num_meshes);
    const auto array_of = [&](const View &v, size_t element_bytes, std::vector<uint8_t> &copy, const std::string &where, size_t &count) {
        if (v.bytes % element_bytes) {
            throw std::runtime_error("crts: " + where + " is not a whole number of elements");
        }
        count = v.bytes / element_bytes;
        if (reinterpret_cast<uintptr_t>(v.data) % 4 == 0) {
            return v.data;
        }
        copy.assign(v.data, v.data + v.bytes);  // (vector storage is aligned)
        return static_cast<const uint8_t *>(copy.data());
    };
    S.unaligned_copies.resize(num_meshes * 3);
    for (size_t i = 0; i < num_meshes; ++i) {
        const std::string where = "mesh " + std::to_string(i);
        const Json &m = meshes.at(i, "meshes");
        size_t nv = 0, nt = 0, nuv = 0;
        const uint8_t *pos = array_of(view_of(m.at("positions", where), where), 12, S.unaligned_copies[3 * i], where + " positions", nv);
        const uint8_t *idx = array_of(view_of(m.at("indices", where), where), 12, S.unaligned_copies[3 * i + 1], where + " indices", nt);
        const uint8_t *uvs = nullptr;
        if (m.kind == Json::kObject && m.find("texcoords")) {
            uvs = array_of(view_of(*m.find("texcoords"), where), 8, S.unaligned_copies[3 * i + 2], where + " texcoords", nuv);
            if (nuv == 0) {
                uvs = nullptr;  // (Geometry::uvs empty = no texture coordinates)
            } else if (nuv != nv) {
                throw std::runtime_error("crts: " + where + " has " + std::to_string(nuv) + " texcoords for " + std::to_string(nv) + " positions");
            }
        }
        if (nv > 0xffffffffull || nt > 0xffffffffull) {
            throw std::runtime_error("crts: " + where + " is too large for 32-bit counts");
        }
        S.geometry_views[i] = crt_geometry_t{reinterpret_cast<const float *>(pos), reinterpret_cast<const float *>(uvs),
                                             reinterpret_cast<const uint32_t *>(idx), (uint32_t)nv, (uint32_t)nt};
        S.meshes[i] = crt_mesh_t{&S.geometry_views[i], 1u};
    }
    const double t_parsed = now_s();
    // ---- images: embedded files, decoded like stbi_load_from_memory(..., 4) with the vertical flip (:488-511) ----
    const Json &images = section("images");
    const size_t num_images = images.size();
    S.texture_data.resize(num_images);
    S.textures.resize(num_images);
    std::vector<View> image_views(num_images);
    std::vector<std::string> image_names(num_images);
    for (size_t i = 0; i < num_images; ++i) {
        const std::string where = "image " + std::to_string(i);
        const Json &img = images.at(i, "images");
        image_views[i] = view_of(img.at("view", where), where);
        image_names[i] = img.at("name", where).string(where + " name");
        S.texture_names.push_back(image_names[i]);
        const int32_t cs = img.at("color_space", where).string(where + " color_space") == "LINEAR" ? CRT_COLOR_SPACE_LINEAR : CRT_COLOR_SPACE_SRGB;
        S.textures[i] = crt_image_t{nullptr, 0, 0, 4, cs};
    }
    std::vector<std::string> image_errors(num_images);
    crt::parallel_blocks((uint32_t)num_images, nthreads, [&](uint32_t i) {
        try {
            int w = 0, h = 0;
            decode_image_rgba(image_views[i].data, image_views[i].bytes, image_names[i], S.texture_data[i], w, h, true);
            S.textures[i].data = S.texture_data[i].data();
            S.textures[i].width = w;
            S.textures[i].height = h;
        } catch (const std::exception &e) {
            image_errors[i] = std::string("Failed to load ") + image_names[i] + " (" + e.what() + ")";
        }
    });
    for (const std::string &e : image_errors) {
        if (!e.empty()) {
            throw std::runtime_error(e);
        }
    }
    // ---- materials (:513-556) ----
    const Json &materials = S.white_diffuse ? null_json : section("materials");  // (:513)
    for (size_t i = 0; i < materials.size(); ++i) {
        const std::string where = "material " + std::to_string(i);
        const Json &m = materials.at(i, "materials");
        crt_material_t d;
        std::memset(&d, 0, sizeof(d));
        const std::vector<float> base = m.at("base_color", where).floats(3, where + " base_color");
        d.base_color[0] = base[0], d.base_color[1] = base[1], d.base_color[2] = base[2];
        if (const Json *t = m.find("base_color_texture")) {
            const uint32_t mask = 0x80000000u | ((uint32_t)t->number<int32_t>(where + " base_color_texture") & 0x1fffffffu);
            std::memcpy(&d.base_color[0], &mask, 4);  // TEXTURED_PARAM_MASK, SET_TEXTURE_ID
        }
        const auto param = [&](const char *name, float &val) {
            val = m.at(name, where).number<float>(where + " " + name);
            const std::string tex_name = std::string(name) + "_texture";
            if (const Json *t = m.find(tex_name.c_str())) {
                const uint32_t id = (uint32_t)t->at("texture", where + " " + tex_name).number<int32_t>(where + " " + tex_name);
                const uint32_t channel = t->at("channel", where + " " + tex_name).number<uint32_t>(where + " " + tex_name);
                const uint32_t mask = 0x80000000u | (id & 0x1fffffffu) | ((channel & 0x3u) << 29);  // SET_TEXTURE_ID, SET_TEXTURE_CHANNEL
                std::memcpy(&val, &mask, 4);
            }
        };
        param("metallic", d.metallic);
        param("specular", d.specular);
        param("roughness", d.roughness);
        param("specular_tint", d.specular_tint);
        param("anisotropic", d.anisotropy);
        param("sheen", d.sheen);
        param("sheen_tint", d.sheen_tint);
        param("clearcoat", d.clearcoat);
        param("clearcoat_roughness", d.clearcoat_gloss);
        param("ior", d.ior);
        param("transmission", d.specular_transmission);
        S.materials.push_back(d);
    }
    // ---- objects (:558-606): instances of (mesh, material) pairs, quad lights, cameras ----
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> pair_ids;  // (mesh, material) -> parameterized mesh, in order of first use
    const Json &objects = section("objects");
    for (size_t i = 0; i < objects.size(); ++i) {
        const std::string where = "object " + std::to_string(i);
        const Json &n = objects.at(i, "objects");
        const std::string &type = n.at("type", where).string(where + " type");
        const std::vector<float> mat = n.at("matrix", where).floats(16, where + " matrix");  // column major (glm::make_mat4)
        const float *col[4] = {&mat[0], &mat[4], &mat[8], &mat[12]};
        if (type == "MESH") {
            const uint64_t mesh_id = n.at("mesh", where).number<uint64_t>(where + " mesh");
            const uint32_t mat_id = S.white_diffuse ? 0xffffffffu : n.at("material", where).number<uint32_t>(where + " material");  // (:568-571)
            if (mesh_id >= num_meshes) {
                throw std::runtime_error("crts: " + where + " instances mesh " + std::to_string(mesh_id) + " of " + std::to_string(num_meshes));
            }
            const auto key = std::make_pair((uint32_t)mesh_id, mat_id);
            auto it = pair_ids.find(key);
            if (it == pair_ids.end()) {
                it = pair_ids.insert(std::make_pair(key, (uint32_t)S.parameterized_meshes.size())).first;
                S.parameterized_meshes.push_back(crt_parameterized_mesh_t{nullptr, 0u, (uint32_t)mesh_id});
                S.material_ids.push_back(std::vector<uint32_t>{mat_id});
            }
            crt_instance_t inst;
            std::memset(&inst, 0, sizeof(inst));
            std::memcpy(inst.transform, mat.data(), sizeof(inst.transform));
            inst.parameterized_mesh_id = it->second;
            S.instances.push_back(inst);
        } else if (type == "LIGHT") {
            crt_quad_light_t L;
            std::memset(&L, 0, sizeof(L));
            const std::vector<float> color = n.at("color", where).floats(3, where + " color");
            const float energy = n.at("energy", where).number<float>(where + " energy");
            for (int k = 0; k < 3; ++k) {
                L.emission[k] = color[k] * energy;
            }
            L.emission[3] = 1.f;
            float unit[4];
            std::memcpy(L.position, col[3], sizeof(L.position));
            normalize4(col[2], unit);
            for (int k = 0; k < 4; ++k) {
                L.normal[k] = -unit[k];
            }
            normalize4(col[0], unit);
            std::memcpy(L.v_x, unit, sizeof(L.v_x));
            normalize4(col[1], unit);
            std::memcpy(L.v_y, unit, sizeof(L.v_y));
            const Json &size = n.at("size", where);
            L.width = size.at((size_t)0, where + " size").number<float>(where + " size");
            L.height = size.at((size_t)1, where + " size").number<float>(where + " size");
            S.lights.push_back(L);
        } else if (type == "CAMERA") {
            crtio_camera_t cam;
            const float back[4] = {-col[2][0], -col[2][1], -col[2][2], -col[2][3]};
            float dir[4], up[4];
            normalize4(back, dir);
            normalize4(col[1], up);
            for (int k = 0; k < 3; ++k) {
                cam.position[k] = col[3][k];
                cam.center[k] = cam.position[k] + dir[k] * 10.f;
                cam.up[k] = up[k];
            }
            cam.fov_y = n.at("fov_y", where).number<float>(where + " fov_y") / 1.18f;
            S.cameras.push_back(cam);
        } else {
            throw std::runtime_error("Unsupported object type: not a mesh or camera?");
        }
    }
    std::ostringstream warn;
    validate_materials(S, warn);
    if (S.lights.empty()) {
        warn << "No lights found in scene, generating one\n";
        S.lights.push_back(generated_light(10.f));  // :612-624
    }
    S.finish_views();
    S.warnings = warn.str();
    const double t_end = now_s();
    S.timings[0] = t_end - t_start;
    S.timings[1] = t_parsed - t_start;
    S.timings[2] = 0.0;
    S.timings[3] = t_end - t_parsed;
}

std::string file_extension_of(const std::string &path)  // get_file_extension, util/util.cpp
{
    const size_t dot = path.rfind('.');
    return dot == std::string::npos ? std::string() : path.substr(dot + 1);
}

// ---------------------------------------------------------------------------------------------------------------
// glTF 2.0 (.gltf + .bin / data: URIs, .glb) as Scene::load_gltf reads it through tinygltf (util/scene.cpp:230-415,
// util/flatten_gltf.cpp). The float arithmetic of the node transforms (quaternion -> matrix, matrix products of the scene
// graph flattening) is glm's, in its order of operations.
struct Mat4 {
    float c[4][4];  // column major: c[column][row]
};
Mat4 mat4_identity()
{
    Mat4 m;
    std::memset(&m, 0, sizeof(m));
    m.c[0][0] = m.c[1][1] = m.c[2][2] = m.c[3][3] = 1.f;
    return m;
}
// glm: each column of the product is a * b[col], a row's four products summed left to right
Mat4 mat4_mul(const Mat4 &a, const Mat4 &b)
{
    Mat4 r;
    for (int col = 0; col < 4; ++col) {
        for (int row = 0; row < 4; ++row) {
            float sum = a.c[0][row] * b.c[col][0];
            for (int k = 1; k < 4; ++k) {
                sum += a.c[k][row] * b.c[col][k];
            }
            r.c[col][row] = sum;
        }
    }
    return r;
}
std::vector<double> json_doubles(const Json &j, const std::string &where)
{
    if (j.kind != Json::kArray) {
        throw std::runtime_error("scene header: " + where + " is not an array of numbers");
    }
    std::vector<double> out(j.items.size());
    for (size_t k = 0; k < out.size(); ++k) {
        out[k] = j.items[k].number<double>(where);
    }
    return out;
}
// read_node_transform (util/flatten_gltf.cpp:9-31): matrix, or translate * rotate * scale
Mat4 gltf_node_transform(const Json &node, const std::string &where)
{
    Mat4 t = mat4_identity();
    const Json *matrix = node.find("matrix");
    if (matrix && matrix->kind == Json::kArray && !matrix->items.empty()) {  // (tinygltf: matrix and T/R/S are exclusive)
        const std::vector<double> m = json_doubles(*matrix, where + " matrix");
        if (m.size() < 16) {
            throw std::runtime_error("glTF: " + where + " has a matrix of " + std::to_string(m.size()) + " numbers");
        }
        for (int k = 0; k < 16; ++k) {
            t.c[k / 4][k % 4] = (float)m[k];
        }
        return t;
    }
    const auto triple = [&](const char *key, size_t n) {
        const Json *j = node.find(key);
        std::vector<double> v;
        if (j && j->kind == Json::kArray && !j->items.empty()) {
            v = json_doubles(*j, where + " " + key);
            if (v.size() < n) {
                throw std::runtime_error("glTF: " + where + " has a " + key + " of " + std::to_string(v.size()) + " numbers");
            }
        }
        return v;
    };
    const std::vector<double> scale = triple("scale", 3), rotation = triple("rotation", 4), translation = triple("translation", 3);
    if (!scale.empty()) {  // glm::scale(v): the identity's columns times v
        const float v[3] = {(float)scale[0], (float)scale[1], (float)scale[2]};
        const Mat4 id = mat4_identity();
        for (int col = 0; col < 3; ++col) {
            for (int row = 0; row < 4; ++row) {
                t.c[col][row] = id.c[col][row] * v[col];
            }
        }
    }
    if (!rotation.empty()) {  // glm::mat4_cast(quat(w = r[3], x = r[0], y = r[1], z = r[2])) * transform
        const float x = (float)rotation[0], y = (float)rotation[1], z = (float)rotation[2], w = (float)rotation[3];
        const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
        Mat4 r = mat4_identity();
        r.c[0][0] = 1.f - 2.f * (qyy + qzz);
        r.c[0][1] = 2.f * (qxy + qwz);
        r.c[0][2] = 2.f * (qxz - qwy);
        r.c[1][0] = 2.f * (qxy - qwz);
        r.c[1][1] = 1.f - 2.f * (qxx + qzz);
        r.c[1][2] = 2.f * (qyz + qwx);
        r.c[2][0] = 2.f * (qxz + qwy);
        r.c[2][1] = 2.f * (qyz - qwx);
        r.c[2][2] = 1.f - 2.f * (qxx + qyy);
        t = mat4_mul(r, t);
    }
    if (!translation.empty()) {  // glm::translate(v): column 3 = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] of the identity
        const float v[3] = {(float)translation[0], (float)translation[1], (float)translation[2]};
        Mat4 tr = mat4_identity();
        const Mat4 id = mat4_identity();
        for (int row = 0; row < 4; ++row) {
            tr.c[3][row] = id.c[0][row] * v[0] + id.c[1][row] * v[1] + id.c[2][row] * v[2] + id.c[3][row];
        }
        t = mat4_mul(tr, t);
    }
    return t;
}

std::string url_decode(const std::string &str)  // dlib::urldecode as vendored in tiny_gltf.h:2235-2255
{
    const auto hex = [](unsigned char ch) -> unsigned char {
        if (ch <= '9' && ch >= '0') {
            return (unsigned char)(ch - '0');
        }
        if (ch <= 'f' && ch >= 'a') {
            return (unsigned char)(ch - 'a' + 10);
        }
        if (ch <= 'F' && ch >= 'A') {
            return (unsigned char)(ch - 'A' + 10);
        }
        return 0;
    };
    std::string out;
    for (size_t i = 0; i < str.size(); ++i) {
        if (str[i] == '+') {
            out += ' ';
        } else if (str[i] == '%' && str.size() > i + 2) {
            out += (char)(unsigned char)((hex((unsigned char)str[i + 1]) << 4) | hex((unsigned char)str[i + 2]));
            i += 2;
        } else {
            out += str[i];
        }
    }
    return out;
}

// The payload of a data: URI with one of the media types tinygltf accepts (IsDataURI, tiny_gltf.h:2820-2857), else false
bool decode_data_uri(const std::string &uri, std::vector<uint8_t> &out)
{
    static const char *const headers[] = {"data:application/octet-stream;base64,", "data:image/jpeg;base64,", "data:image/png;base64,",
                                          "data:image/bmp;base64,",  "data:image/gif;base64,", "data:text/plain;base64,",
                                          "data:application/gltf-buffer;base64,"};
    size_t start = 0;
    for (const char *h : headers) {
        if (uri.compare(0, std::strlen(h), h) == 0) {
            start = std::strlen(h);
        }
    }
    if (!start) {
        return false;
    }
    out.clear();
    out.reserve((uri.size() - start) / 4 * 3);
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = start; i < uri.size(); ++i) {
        const char ch = uri[i];
        int v;
        if (ch >= 'A' && ch <= 'Z') {
            v = ch - 'A';
        } else if (ch >= 'a' && ch <= 'z') {
            v = ch - 'a' + 26;
        } else if (ch >= '0' && ch <= '9') {
            v = ch - '0' + 52;
        } else if (ch == '+') {
            v = 62;
        } else if (ch == '/') {
            v = 63;
        } else {
            break;  // '=' padding (or anything that is not base64) ends the data
        }
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return true;
}

void load_gltf_impl(const std::string &file, int threads, crtio_scene &S)
{
    const double t_start = now_s();
    const unsigned nthreads = crt::host_threads(threads);
    S.mapping.reset(new MappedFile(file, /*sequential=*/false));
    const MappedFile &map = *S.mapping;
    const bool binary = file_extension_of(file) != "gltf";
    const char *json_begin = map.data, *json_end = map.data + map.size;
    const uint8_t *glb_bin = nullptr;
    size_t glb_bin_size = 0;
    if (binary) {  // TinyGLTF::LoadBinaryFromMemory (tiny_gltf.h:6194-6260)
        if (map.size < 20) {
            throw std::runtime_error("TinyGLTF Error loading " + file + " error: Too short data size for glTF Binary.");
        }
        if (std::memcmp(map.data, "glTF", 4) != 0) {
            throw std::runtime_error("TinyGLTF Error loading " + file + " error: Invalid magic.");
        }
        uint32_t length, model_length, model_format;
        std::memcpy(&length, map.data + 8, 4);
        std::memcpy(&model_length, map.data + 12, 4);
        std::memcpy(&model_format, map.data + 16, 4);
        if (20ull + model_length > map.size || model_length < 1 || length > map.size || 20ull + model_length > length ||
            model_format != 0x4E4F534Au) {
            throw std::runtime_error("TinyGLTF Error loading " + file + " error: Invalid glTF binary.");
        }
        json_begin = map.data + 20;
        json_end = json_begin + model_length;
        const size_t rest = (size_t)length - (20 + (size_t)model_length);
        if (rest >= 8) {  // the BIN chunk: 4 bytes length, 4 bytes type, data
            glb_bin = reinterpret_cast<const uint8_t *>(json_end) + 8;
            glb_bin_size = rest - 8;
        }
    }
    const Json doc = JsonParser(json_begin, json_end).parse_document();
    if (doc.kind != Json::kObject) {
        throw std::runtime_error("TinyGLTF Error loading " + file + " error: the document is not a JSON object");
    }
    static const Json null_json;
    const auto section = [&](const char *key) -> const Json & {
        const Json *j = doc.find(key);
        return j ? *j : null_json;
    };
    const auto optional_index = [&](const Json &o, const char *key, const std::string &where) -> int64_t {
        const Json *j = o.kind == Json::kObject ? o.find(key) : nullptr;
        if (!j) {
            return -1;
        }
        if (j->kind != Json::kUnsigned && j->kind != Json::kSigned) {
            throw std::runtime_error("glTF: " + where + " " + key + " is not an integer");
        }
        return j->number<int64_t>(where);
    };
    const auto required_index = [&](const Json &o, const char *key, const std::string &where) -> size_t {
        const int64_t v = optional_index(o, key, where);
        if (v < 0) {
            throw std::runtime_error("glTF: " + where + " has no \"" + key + "\"");
        }
        return (size_t)v;
    };
    std::string base_dir;  // GetBaseDir (tiny_gltf.h): up to the last separator, "" if there is none
    {
        const size_t sep = file.find_last_of("/\\");
        base_dir = sep == std::string::npos ? std::string() : file.substr(0, sep);
    }
    const auto join_path = [&](const std::string &name) {
        if (base_dir.empty()) {
            return name;
        }
        return base_dir.back() == '/' ? base_dir + name : base_dir + "/" + name;
    };
    // ---- buffers ----
    struct Span {
        const uint8_t *data = nullptr;
        size_t size = 0;
    };
    const Json &jbuffers = section("buffers");
    std::vector<Span> buffers(jbuffers.size());
    for (size_t i = 0; i < buffers.size(); ++i) {
        const std::string where = "buffer " + std::to_string(i);
        const Json &b = jbuffers.at(i, "buffers");
        const size_t byte_length = required_index(b, "byteLength", where);
        const Json *uri = b.find("uri");
        if (!uri) {
            if (!binary || i != 0 || !glb_bin) {
                throw std::runtime_error("glTF: " + where + " has no uri (only the first buffer of a .glb may)");
            }
            buffers[i] = Span{glb_bin, glb_bin_size};
        } else {
            const std::string &u = uri->string(where + " uri");
            S.decoded_buffers.emplace_back();
            if (decode_data_uri(u, S.decoded_buffers.back())) {
                buffers[i] = Span{S.decoded_buffers.back().data(), S.decoded_buffers.back().size()};
            } else {
                S.decoded_buffers.pop_back();
                S.buffer_mappings.emplace_back(new MappedFile(join_path(url_decode(u)), /*sequential=*/false));
                buffers[i] = Span{reinterpret_cast<const uint8_t *>(S.buffer_mappings.back()->data), S.buffer_mappings.back()->size};
            }
        }
        if (buffers[i].size < byte_length) {
            throw std::runtime_error("glTF: " + where + " holds " + std::to_string(buffers[i].size) + " bytes, byteLength says " +
                                     std::to_string(byte_length));
        }
    }
    // ---- buffer views and accessors ----
    const Json &jviews = section("bufferViews");
    const Json &jaccessors = section("accessors");
    struct ViewSpan {
        const uint8_t *data;
        size_t length, stride, room;  // room: bytes from data to the end of the buffer
    };
    const auto view_of = [&](size_t id, const std::string &where) {
        const std::string vw = where + " bufferView " + std::to_string(id);
        const Json &v = jviews.at(id, "bufferViews");
        const size_t buffer = required_index(v, "buffer", vw), length = required_index(v, "byteLength", vw);
        const int64_t offset = std::max<int64_t>(0, optional_index(v, "byteOffset", vw)), stride = std::max<int64_t>(0, optional_index(v, "byteStride", vw));
        if (buffer >= buffers.size()) {
            throw std::runtime_error("glTF: " + vw + " names buffer " + std::to_string(buffer) + " of " + std::to_string(buffers.size()));
        }
        if ((size_t)offset > buffers[buffer].size || length > buffers[buffer].size - (size_t)offset) {
            throw std::runtime_error("glTF: " + vw + " reaches past the end of its buffer");
        }
        return ViewSpan{buffers[buffer].data + offset, length, (size_t)stride, buffers[buffer].size - (size_t)offset};
    };
    struct Elements {  // Accessor<T> (util/buffer_view.h): element i at data + i * stride
        const uint8_t *data;
        size_t stride, count;
        int component_type;
        std::string type;
    };
    const auto accessor_of = [&](size_t id, size_t element_bytes, const std::string &where) {
        const std::string aw = where + " accessor " + std::to_string(id);
        const Json &a = jaccessors.at(id, "accessors");
        if (a.find("sparse")) {
            throw std::runtime_error("glTF: " + aw + " is sparse (not supported)");
        }
        const int64_t view_id = optional_index(a, "bufferView", aw);
        if (view_id < 0) {
            throw std::runtime_error("glTF: " + aw + " has no bufferView");
        }
        const ViewSpan view = view_of((size_t)view_id, aw);
        Elements e;
        e.component_type = (int)required_index(a, "componentType", aw);
        e.type = a.at("type", aw).string(aw + " type");
        e.count = required_index(a, "count", aw);
        size_t components = 0, component_bytes = 0;
        static const struct {
            const char *name;
            size_t n;
        } types[] = {{"SCALAR", 1}, {"VEC2", 2}, {"VEC3", 3}, {"VEC4", 4}, {"MAT2", 4}, {"MAT3", 9}, {"MAT4", 16}};
        for (const auto &t : types) {
            components = e.type == t.name ? t.n : components;
        }
        switch (e.component_type) {
        case 5120: case 5121: component_bytes = 1; break;
        case 5122: case 5123: component_bytes = 2; break;
        case 5124: case 5125: case 5126: component_bytes = 4; break;
        case 5130: component_bytes = 8; break;
        default: break;
        }
        if (!components || !component_bytes) {
            throw std::runtime_error("glTF: " + aw + " has an unknown type or componentType");
        }
        e.stride = std::max(view.stride, components * component_bytes);  // BufferView: max(byteStride, gltf_base_stride)
        const size_t offset = (size_t)std::max<int64_t>(0, optional_index(a, "byteOffset", aw));
        if (e.count && (offset > view.room || (e.count - 1) * e.stride + element_bytes > view.room - offset)) {
            throw std::runtime_error("glTF: " + aw + " reaches past the end of its buffer");
        }
        e.data = view.data + offset;
        return e;
    };
    // elements of `bytes` bytes each as one packed, 4-byte aligned array: in place if they already are, else gathered
    const auto packed = [&](const Elements &e, size_t bytes) {
        if (e.stride == bytes && reinterpret_cast<uintptr_t>(e.data) % 4 == 0) {
            return e.data;
        }
        S.unaligned_copies.emplace_back(e.count * bytes);
        uint8_t *dst = S.unaligned_copies.back().data();
        for (size_t i = 0; i < e.count; ++i) {
            std::memcpy(dst + i * bytes, e.data + i * e.stride, bytes);
        }
        return static_cast<const uint8_t *>(dst);
    };
    // ---- meshes: a glTF mesh is a parameterized mesh, its primitives are the geometries (scene.cpp:256-330) ----
    const Json &jmeshes = section("meshes");
    const size_t num_meshes = jmeshes.size();
    size_t num_geometries = 0;
    for (size_t m = 0; m < num_meshes; ++m) {
        num_geometries += jmeshes.at(m, "meshes").at("primitives", "mesh " + std::to_string(m)).size();
    }
    S.geometry_views.reserve(num_geometries);  // (crt_mesh_t points into it)
    S.meshes.resize(num_meshes);
    S.material_ids.resize(num_meshes);
    S.parameterized_meshes.resize(num_meshes);
    for (size_t m = 0; m < num_meshes; ++m) {
        const Json &prims = jmeshes.at(m, "meshes").at("primitives", "mesh " + std::to_string(m));
        const size_t first = S.geometry_views.size();
        for (size_t k = 0; k < prims.size(); ++k) {
            const std::string where = "mesh " + std::to_string(m) + " primitive " + std::to_string(k);
            const Json &p = prims.at(k, where);
            S.material_ids[m].push_back(S.white_diffuse ? 0xffffffffu : (uint32_t)optional_index(p, "material", where));  // -1: validate_materials
            const int64_t mode = optional_index(p, "mode", where);
            if (mode != -1 && mode != 4) {
                throw std::runtime_error("Unsupported primitive mode! Only triangles are supported");
            }
            const Json &attributes = p.at("attributes", where);
            const Elements pos = accessor_of(required_index(attributes, "POSITION", where), 12, where + " POSITION");
            if (pos.component_type != 5126 || pos.type != "VEC3") {
                throw std::runtime_error("glTF: " + where + " POSITION is not FLOAT VEC3");
            }
            const uint8_t *uv_data = nullptr;
            if (attributes.find("TEXCOORD_0")) {
                const Elements uv = accessor_of(required_index(attributes, "TEXCOORD_0", where), 8, where + " TEXCOORD_0");
                if (uv.component_type != 5126 || uv.type != "VEC2") {
                    throw std::runtime_error("glTF: " + where + " TEXCOORD_0 is not FLOAT VEC2 (the reference reads it as floats)");
                }
                if (uv.count != pos.count) {
                    throw std::runtime_error("glTF: " + where + " has " + std::to_string(uv.count) + " texcoords for " +
                                             std::to_string(pos.count) + " positions");
                }
                uv_data = uv.count ? packed(uv, 8) : nullptr;
            }
            const int64_t indices_id = optional_index(p, "indices", where);
            if (indices_id < 0) {
                throw std::runtime_error("glTF: " + where + " has no indices (the reference reads accessor -1 there)");
            }
            const std::string iw = where + " indices";
            const Json &ia = jaccessors.at((size_t)indices_id, "accessors");
            const int component_type = (int)required_index(ia, "componentType", iw);
            const uint8_t *index_data = nullptr;
            size_t num_tris = 0;
            if (component_type == 5123) {  // uint16 -> uint32
                const Elements idx = accessor_of((size_t)indices_id, 2, iw);
                num_tris = idx.count / 3;
                S.unaligned_copies.emplace_back(num_tris * 12);
                uint32_t *dst = reinterpret_cast<uint32_t *>(S.unaligned_copies.back().data());
                for (size_t i = 0; i < num_tris * 3; ++i) {
                    uint16_t v;
                    std::memcpy(&v, idx.data + i * idx.stride, 2);
                    dst[i] = v;
                }
                index_data = S.unaligned_copies.back().data();
            } else if (component_type == 5125) {
                Elements idx = accessor_of((size_t)indices_id, 4, iw);
                num_tris = idx.count / 3;
                idx.count = num_tris * 3;
                index_data = packed(idx, 4);
            } else {
                throw std::runtime_error("Unsupported index component type");
            }
            if (pos.count > 0xffffffffull || num_tris > 0xffffffffull) {
                throw std::runtime_error("glTF: " + where + " is too large for 32-bit counts");
            }
            S.geometry_views.push_back(crt_geometry_t{reinterpret_cast<const float *>(packed(pos, 12)), reinterpret_cast<const float *>(uv_data),
                                                      reinterpret_cast<const uint32_t *>(index_data), (uint32_t)pos.count, (uint32_t)num_tris});
        }
        S.meshes[m] = crt_mesh_t{S.geometry_views.data() + first, (uint32_t)prims.size()};
        S.parameterized_meshes[m] = crt_parameterized_mesh_t{nullptr, 0u, (uint32_t)m};
    }
    const double t_parsed = now_s();
    // ---- images (tinygltf's LoadImageData: stb_image, four components, rows top-down), linear until a material says otherwise ----
    const Json &jimages = S.white_diffuse ? null_json : section("images");  // (images and materials: scene.cpp:329)
    const size_t num_images = jimages.size();
    S.texture_data.resize(num_images);
    S.textures.resize(num_images);
    std::vector<Span> encoded(num_images);
    std::vector<std::string> image_names(num_images);
    std::vector<std::vector<uint8_t>> image_files(num_images);
    for (size_t i = 0; i < num_images; ++i) {
        const std::string where = "image " + std::to_string(i);
        const Json &img = jimages.at(i, "images");
        if (const Json *name = img.find("name")) {
            image_names[i] = name->kind == Json::kString ? name->str : std::string();
        }
        S.texture_names.push_back(image_names[i]);
        image_names[i] = where + " \"" + image_names[i] + "\"";
        const Json *uri = img.find("uri");
        const int64_t view_id = optional_index(img, "bufferView", where);
        if ((uri != nullptr) == (view_id >= 0)) {
            throw std::runtime_error("TinyGLTF Error loading " + file + " error: " + where + " needs exactly one of `bufferView` and `uri`");
        }
        if (view_id >= 0) {
            const ViewSpan v = view_of((size_t)view_id, where);
            encoded[i] = Span{v.data, v.length};
        } else {
            const std::string &u = uri->string(where + " uri");
            if (!decode_data_uri(u, image_files[i])) {
                const std::string path = join_path(url_decode(u));
                std::ifstream in(path.c_str(), std::ios::binary);
                if (!in) {
                    throw std::runtime_error("glTF: cannot read " + path + " (" + where + ")");
                }
                image_files[i].assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
            }
            encoded[i] = Span{image_files[i].data(), image_files[i].size()};
        }
        S.textures[i] = crt_image_t{nullptr, 0, 0, 4, CRT_COLOR_SPACE_LINEAR};
    }
    std::vector<std::string> image_errors(num_images);
    crt::parallel_blocks((uint32_t)num_images, nthreads, [&](uint32_t i) {
        try {
            int w = 0, h = 0;
            // (tinygltf keeps the 16 bits of a 16-bit PNG, and Scene::load_gltf refuses such an image: scene.cpp:335-338)
            if (encoded[i].size > 24 && std::memcmp(encoded[i].data, "\x89PNG", 4) == 0 && encoded[i].data[24] == 16) {
                throw std::runtime_error("Unsupported image pixel type");
            }
            decode_image_rgba(encoded[i].data, encoded[i].size, image_names[i], S.texture_data[i], w, h, false);
            S.textures[i].data = S.texture_data[i].data();
            S.textures[i].width = w;
            S.textures[i].height = h;
        } catch (const std::exception &e) {
            image_errors[i] = e.what();
        }
    });
    for (const std::string &e : image_errors) {
        if (!e.empty()) {
            throw std::runtime_error(e);
        }
    }
    // ---- materials: pbrMetallicRoughness -> DisneyMaterial (scene.cpp:353-388) ----
    const Json &jtextures = section("textures");
    const auto image_of_texture = [&](const Json &info, const std::string &where) {
        const size_t tex = required_index(info, "index", where);
        const size_t source = required_index(jtextures.at(tex, "textures"), "source", where + " texture " + std::to_string(tex));
        if (source >= num_images) {
            throw std::runtime_error("glTF: " + where + " names image " + std::to_string(source) + " of " + std::to_string(num_images));
        }
        return (uint32_t)source;
    };
    const Json &jmaterials = S.white_diffuse ? null_json : section("materials");
    for (size_t i = 0; i < jmaterials.size(); ++i) {
        const std::string where = "material " + std::to_string(i);
        const Json &m = jmaterials.at(i, "materials");
        crt_material_t d;  // DisneyMaterial's defaults, util/material.h:29-46
        std::memset(&d, 0, sizeof(d));
        d.ior = 1.5f;
        double base[3] = {1.0, 1.0, 1.0}, metallic = 1.0, roughness = 1.0;  // tinygltf's PbrMetallicRoughness defaults
        const Json *pbr = m.kind == Json::kObject ? m.find("pbrMetallicRoughness") : nullptr;
        if (pbr && pbr->kind == Json::kObject) {
            if (const Json *f = pbr->find("baseColorFactor")) {
                const std::vector<double> v = json_doubles(*f, where + " baseColorFactor");
                if (v.size() != 4) {
                    throw std::runtime_error("TinyGLTF Error loading " + file + " error: Array length of `baseColorFactor` parameter in "
                                             "pbrMetallicRoughness must be 4, but got " + std::to_string(v.size()));
                }
                base[0] = v[0], base[1] = v[1], base[2] = v[2];
            }
            if (const Json *f = pbr->find("metallicFactor")) {
                metallic = f->number<double>(where + " metallicFactor");
            }
            if (const Json *f = pbr->find("roughnessFactor")) {
                roughness = f->number<double>(where + " roughnessFactor");
            }
        }
        d.base_color[0] = (float)base[0], d.base_color[1] = (float)base[1], d.base_color[2] = (float)base[2];
        d.metallic = (float)metallic;
        d.roughness = (float)roughness;
        if (pbr && pbr->kind == Json::kObject) {
            if (const Json *t = pbr->find("baseColorTexture")) {
                const uint32_t id = image_of_texture(*t, where + " baseColorTexture");
                S.textures[id].color_space = CRT_COLOR_SPACE_SRGB;
                const uint32_t mask = 0x80000000u | (id & 0x1fffffffu);
                std::memcpy(&d.base_color[0], &mask, 4);
            }
            if (const Json *t = pbr->find("metallicRoughnessTexture")) {  // metallic = blue, roughness = green
                const uint32_t id = image_of_texture(*t, where + " metallicRoughnessTexture");
                S.textures[id].color_space = CRT_COLOR_SPACE_LINEAR;
                const uint32_t metal = 0x80000000u | (id & 0x1fffffffu) | (2u << 29), rough = 0x80000000u | (id & 0x1fffffffu) | (1u << 29);
                std::memcpy(&d.metallic, &metal, 4);
                std::memcpy(&d.roughness, &rough, 4);
            }
        }
        S.materials.push_back(d);
    }
    // ---- instances: the nodes of the default scene that carry a mesh, the scene graph flattened (flatten_gltf.cpp) ----
    const Json &jscenes = section("scenes"), &jnodes = section("nodes");
    int64_t default_scene = optional_index(doc, "scene", "document");
    if (default_scene < 0) {
        default_scene = 0;
    }
    const Json &scene = jscenes.at((size_t)default_scene, "scenes");
    const Json *roots = scene.find("nodes");
    std::vector<size_t> root_ids;
    for (size_t k = 0; roots && k < roots->size(); ++k) {
        root_ids.push_back(roots->at(k, "scene nodes").number<size_t>("scene nodes"));
    }
    const auto node_at = [&](size_t id) -> const Json & { return jnodes.at(id, "nodes"); };
    const auto add_instance = [&](const Json &node, const Mat4 &transform, const std::string &where) {
        const int64_t mesh = optional_index(node, "mesh", where);
        if (mesh < 0) {
            return;
        }
        if ((size_t)mesh >= num_meshes) {
            throw std::runtime_error("glTF: " + where + " instances mesh " + std::to_string(mesh) + " of " + std::to_string(num_meshes));
        }
        crt_instance_t inst;
        std::memset(&inst, 0, sizeof(inst));
        std::memcpy(inst.transform, transform.c, sizeof(inst.transform));
        inst.parameterized_mesh_id = (uint32_t)mesh;
        S.instances.push_back(inst);
    };
    bool single_level = true;  // gltf_is_single_level: no root has children
    for (size_t id : root_ids) {
        const Json *children = node_at(id).find("children");
        single_level = single_level && !(children && children->size());
    }
    if (single_level) {
        for (size_t id : root_ids) {
            add_instance(node_at(id), gltf_node_transform(node_at(id), "node " + std::to_string(id)), "node " + std::to_string(id));
        }
    } else {
        // depth first, parents before children; a node's matrix = parent * own (the float matrix goes through the flattened
        // node's double array and back unchanged)
        struct Visit {
            size_t id;
            Mat4 parent;
            int depth;
        };
        std::vector<Visit> stack;
        for (size_t k = root_ids.size(); k-- > 0;) {
            stack.push_back(Visit{root_ids[k], mat4_identity(), 0});
        }
        size_t visited = 0;
        while (!stack.empty()) {
            const Visit v = stack.back();
            stack.pop_back();
            if (v.depth > 256 || ++visited > (size_t)1 << 26) {
                throw std::runtime_error("glTF: the node hierarchy of " + file + " is cyclic or too deep");
            }
            const std::string where = "node " + std::to_string(v.id);
            const Json &node = node_at(v.id);
            const Mat4 transform = mat4_mul(v.parent, gltf_node_transform(node, where));
            add_instance(node, transform, where);
            const Json *children = node.find("children");
            for (size_t k = children ? children->size() : 0; k-- > 0;) {
                stack.push_back(Visit{children->at(k, where + " children").number<size_t>(where + " children"), transform, v.depth + 1});
            }
        }
    }
    std::ostringstream warn;
    validate_materials(S, warn);
    S.lights.push_back(generated_light(20.f));  // scene.cpp:403-414
    S.finish_views();
    S.warnings = warn.str();
    const double t_end = now_s();
    S.timings[0] = t_end - t_start;
    S.timings[1] = t_parsed - t_start;
    S.timings[2] = 0.0;
    S.timings[3] = t_end - t_parsed;
}

template <typename Fn>
int load_with(const char *path, crtio_scene **out, const char *api, Fn &&load, int material_mode = CRTIO_MATERIALS_DEFAULT)
{
    try {
        if (!path || !out) {
            throw std::runtime_error(std::string(api) + ": null argument");
        }
        *out = nullptr;
        if (material_mode != CRTIO_MATERIALS_DEFAULT && material_mode != CRTIO_MATERIALS_WHITE_DIFFUSE) {
            throw std::runtime_error(std::string(api) + ": unknown material mode " + std::to_string(material_mode));
        }
        std::unique_ptr<crtio_scene> s(new crtio_scene());
        s->white_diffuse = material_mode == CRTIO_MATERIALS_WHITE_DIFFUSE;
        load(std::string(path), *s);
        *out = s.release();
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    } catch (...) {
        g_last_error = "unknown exception";
        return 1;
    }
}

}  // namespace

extern "C" {

int crtio_load_obj(const char *path, int threads, crtio_scene **out)
{
    return load_with(path, out, "crtio_load_obj", [&](const std::string &file, crtio_scene &s) { load_obj_impl(file, threads, s); });
}

int crtio_load_crts(const char *path, int threads, crtio_scene **out)
{
    return load_with(path, out, "crtio_load_crts", [&](const std::string &file, crtio_scene &s) { load_crts_impl(file, threads, s); });
}

int crtio_load_gltf(const char *path, int threads, crtio_scene **out)
{
    return load_with(path, out, "crtio_load_gltf", [&](const std::string &file, crtio_scene &s) { load_gltf_impl(file, threads, s); });
}

int crtio_load_mode(const char *path, int threads, int material_mode, crtio_scene **out)
{
    return load_with(path, out, "crtio_load", [&](const std::string &file, crtio_scene &s) {
        const std::string ext = file_extension_of(file);
        if (ext == "obj") {
            load_obj_impl(file, threads, s);
        } else if (ext == "gltf" || ext == "glb") {
            load_gltf_impl(file, threads, s);
        } else if (ext == "crts") {
            load_crts_impl(file, threads, s);
        } else {
            throw std::runtime_error("Unsupported file " + file);  // scene.cpp:63-66 (PBRT is not read, as in a default build of the reference)
        }
    }, material_mode);
}

int crtio_load(const char *path, int threads, crtio_scene **out)
{
    return crtio_load_mode(path, threads, CRTIO_MATERIALS_DEFAULT, out);
}

const char *crtio_texture_name(const crtio_scene *s, uint32_t i)
{
    return s && i < s->texture_names.size() ? s->texture_names[i].c_str() : "";
}

int crtio_cameras(const crtio_scene *s, const crtio_camera_t **out)
{
    if (!s) {
        return 0;
    }
    if (out) {
        *out = s->cameras.data();
    }
    return (int)s->cameras.size();
}

const crt_scene_t *crtio_scene_view(const crtio_scene *s)
{
    return s ? &s->view : nullptr;
}

int crtio_timings(const crtio_scene *s, double *out, int n)
{
    if (!s || !out) {
        return 0;
    }
    const int m = std::min(n, 4);
    for (int i = 0; i < m; ++i) {
        out[i] = s->timings[i];
    }
    return m;
}

const char *crtio_warnings(const crtio_scene *s)
{
    return s ? s->warnings.c_str() : "";
}

void crtio_free(crtio_scene *s)
{
    delete s;
}

const char *crtio_last_error(void)
{
    return g_last_error.c_str();
}
}
