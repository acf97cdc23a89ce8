// mesh.cu -- host-side scene ingestion for the accelerated path: Wavefront OBJ -> TriangleGPU[]
// and the model-matrix setters that produce MeshModelGPU.  No device code: these run before the first upload, need no
// rtr_ctx, and report errors through rtr_last_error(NULL).
//
//   rtr_obj_load / rtr_obj_parse   <- cr::Mesh::load             srcCommon/scene/geometry/mesh.cpp:186-263
//   rtr_mesh_primitive             <- cr::Mesh::primitive*       mesh.cpp:64-183
//   rtr_mesh_set_*                 <- cr::Mesh::setPosition/setScale/setRotation/setMaterial/setModel  mesh.cpp:18-62
//   rtr_triangle_centroid          <- cr::Triangle::getCentroid  triangle.cpp:30-32
//
// Mesh::load delegates the parsing to tinyobjloader (system dependency of the reference, CMakeLists.txt:46, not
// vendored and not pinned).  What is restated here is the published behaviour of tinyobjloader 1.2.0 -- the copy inside
// the reference's own tree (srcVulkan/dep/slang/external/tinyobjloader) -- for the part of an OBJ file Mesh::load
// consumes: `v` positions, `f` corner lists (v, v/vt, v//vn, v/vt/vn, negative = relative), grouping statements as
// the points where pending faces are triangulated, and its decimal reader, which is NOT correctly rounded (digit by
// digit accumulation in double, then pow/ldexp for the exponent) and therefore decides the vertex bits.  The triangle
// order of Mesh::load is the file's face order; polygons are split by that version's ear clipping.
// tests/test_mesh_cpu.py checks every triangle against the reference's mesh.cpp compiled with that tinyobjloader
// (oracle/_ref/libref_mesh.so) and against committed fixtures.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "../../include/rtr_scene.hpp"  // cr::Camera: the same code a C++ caller compiles

namespace {

inline bool is_digit(char c) { return static_cast<unsigned>(c - '0') < 10u; }
inline bool is_blank(char c) { return c == ' ' || c == '\t'; }

// Decimal reader of tinyobjloader 1.2.0 (tryParseDouble): [sign] digits [. digits] [(e|E) [sign] digits], greedy,
// trailing garbage ignored.  A token that does not start with a sign or a digit (".5", "nan") is a failure and the
// caller keeps its default.  `end` points at the token's delimiter, which is never a digit.
bool read_decimal(const char* p, const char* end, double* out) {
    if (p >= end) return false;
    int sign = 1;
    if (*p == '+' || *p == '-') {
        if (*p == '-') sign = -1;
        ++p;
    } else if (!is_digit(*p)) {
        return false;
    }
    double mant = 0.0;
    int ndigits = 0;
    for (; p != end && is_digit(*p); ++p, ++ndigits) {
        mant *= 10;
        mant += static_cast<int>(*p - '0');
    }
    if (ndigits == 0) return false;
    int e10 = 0;
    if (p != end && (*p == '.' || *p == 'e' || *p == 'E')) {
        if (*p == '.') {
            static const double tenth[8] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            ++p;
            for (int k = 1; p != end && is_digit(*p); ++p, ++k)
                mant += static_cast<int>(*p - '0') * (k < 8 ? tenth[k] : pow(10.0, static_cast<double>(-k)));
        }
        if (p != end && (*p == 'e' || *p == 'E')) {
            ++p;
            bool eneg = false;
            if (p != end && (*p == '+' || *p == '-')) {
                eneg = *p == '-';
                ++p;
            } else if (p == end || !is_digit(*p)) {
                return false;  // "1e", "1ex"
            }
            int nexp = 0;
            for (; p != end && is_digit(*p); ++p, ++nexp)
                if (e10 < 100000) e10 = e10 * 10 + (*p - '0');  // the original overflows an int here; inf/0 either way
            if (nexp == 0) return false;
            if (eneg) e10 = -e10;
        }
    }
    *out = sign * (e10 ? ldexp(mant * pow(5.0, static_cast<double>(e10)), e10) : mant);
    return true;
}

// one whitespace-delimited real of a `v` statement; missing or unparsable -> 0
float next_real(const char*& p, const char* le) {
    while (p < le && is_blank(*p)) ++p;
    const char* e = p;
    while (e < le && !is_blank(*e)) ++e;
    double val = 0.0;
    read_decimal(p, e, &val);
    p = e;
    return static_cast<float>(val);
}

// atoi() on a bounded buffer: optional white space, optional sign, digits; 0 when there are none
int bounded_atoi(const char* p, const char* le) {
    while (p < le && (*p == ' ' || (*p >= '\t' && *p <= '\r'))) ++p;
    bool neg = false;
    if (p < le && (*p == '+' || *p == '-')) { neg = *p == '-'; ++p; }
    long long acc = 0;
    for (; p < le && is_digit(*p); ++p)
        if (acc < (1ll << 40)) acc = acc * 10 + (*p - '0');
    if (neg) acc = -acc;
    return static_cast<int>(acc);
}

inline const char* skip_to_slash_or_blank(const char* p, const char* le) {
    while (p < le && *p != '/' && !is_blank(*p)) ++p;
    return p;
}

// One corner "v", "v/vt", "v//vn" or "v/vt/vn".  Index 0 (or no number) in any present field is an error, like
// tinyobjloader's fixIndex; only the position index is kept: v > 0 names vertex v - 1 of the file, v < 0 the |v|-th
// last vertex read before this statement.
bool read_corner(const char*& p, const char* le, int* v_raw) {
    const int v = bounded_atoi(p, le);
    if (v == 0) return false;
    *v_raw = v;
    p = skip_to_slash_or_blank(p, le);
    if (p >= le || *p != '/') return true;
    ++p;
    if (p < le && *p == '/') {  // v//vn
        ++p;
        if (bounded_atoi(p, le) == 0) return false;
        p = skip_to_slash_or_blank(p, le);
        return true;
    }
    if (bounded_atoi(p, le) == 0) return false;  // vt
    p = skip_to_slash_or_blank(p, le);
    if (p >= le || *p != '/') return true;
    ++p;
    if (bounded_atoi(p, le) == 0) return false;  // vn
    p = skip_to_slash_or_blank(p, le);
    return true;
}

// What one thread extracts from its piece of the text (whole lines).  Vertex references stay relative to the piece:
// a corner is v - 1 (absolute, >= 0) or kRelative + (vertices of the piece read so far) + v for v < 0, resolved once
// the number of vertices in the pieces before it is known.
constexpr long long kRelative = 1ll << 40;
struct ObjPiece {
    const char* begin = nullptr;
    const char* end = nullptr;
    std::vector<float> pos;            // x y z per `v`
    std::vector<long long> corners;    // see above, faces back to back
    std::vector<uint32_t> face_end;    // end offset into `corners` per face
    struct Flush { uint32_t faces_before, vertices_before; };
    std::vector<Flush> flushes;        // `g` / `o` statements: where pending faces are triangulated
    uint64_t lines = 0;
    uint64_t bad_line = 0;             // 1-based line (within the piece) of the first malformed `f`, 0 = none
};

struct ObjReader {
    std::vector<float> pos;         // x y z per `v`, the whole file
    std::vector<int> corners;       // position indices of the pending faces, back to back
    std::vector<uint32_t> face_end; // end offset into `corners` per pending face
    std::vector<int> tri;           // emitted triangles, three position indices each
};

// crossing-number test of tinyobjloader 1.2.0 (pnpoly) on the 2-D triangle (x[], y[])
bool inside_triangle(const float x[3], const float y[3], float tx, float ty) {
    bool in = false;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((y[i] > ty) != (y[j] > ty)) && (tx < (x[j] - x[i]) * (ty - y[i]) / (y[j] - y[i]) + x[i])) in = !in;
    return in;
}

// Splits one face into triangles the way tinyobjloader 1.2.0 does with `triangulate` on (exportFaceGroupToShape):
// a triangle passes through; a larger polygon is projected on the two axes chosen from its first non-degenerate
// corner and ear-clipped, starting each search where the last ear was cut, giving up after ten trips around the
// polygon (the vertices left over are dropped).  pos[0, nf) is the vertex array AS READ SO FAR: corners that point
// beyond it are treated as that version treats them (skipped in the setup loops, (0,0) in the ear test).
void split_face(const int* idx, size_t n, const float* pos, size_t nf, std::vector<int>& tri) {
    if (n < 3) return;
    if (n == 3) {
        tri.insert(tri.end(), idx, idx + 3);
        return;
    }
    auto ok = [&](int vi, size_t comp) { return static_cast<size_t>(vi) * 3 + comp < nf; };  // negative -> huge -> false
    size_t ax0 = 1, ax1 = 2;
    for (size_t k = 0; k < n; ++k) {
        const int a = idx[k], b = idx[(k + 1) % n], c = idx[(k + 2) % n];
        if (!ok(a, 2) || !ok(b, 2) || !ok(c, 2)) continue;
        const float* A = &pos[size_t(a) * 3];
        const float* B = &pos[size_t(b) * 3];
        const float* C = &pos[size_t(c) * 3];
        const float e0x = B[0] - A[0], e0y = B[1] - A[1], e0z = B[2] - A[2];
        const float e1x = C[0] - B[0], e1y = C[1] - B[1], e1z = C[2] - B[2];
        const float cx = fabsf(e0y * e1z - e0z * e1y);
        const float cy = fabsf(e0z * e1x - e0x * e1z);
        const float cz = fabsf(e0x * e1y - e0y * e1x);
        const float eps = 1.1920928955078125e-07f;  // FLT_EPSILON
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                ax0 = 0;
                if (cz > cx && cz > cy) ax1 = 1;
            }
            break;
        }
    }
    float area = 0.f;
    for (size_t k = 0; k < n; ++k) {
        const int a = idx[k], b = idx[(k + 1) % n];
        if (!ok(a, ax0) || !ok(a, ax1) || !ok(b, ax0) || !ok(b, ax1)) continue;
        const float ax = pos[size_t(a) * 3 + ax0], ay = pos[size_t(a) * 3 + ax1];
        const float bx = pos[size_t(b) * 3 + ax0], by = pos[size_t(b) * 3 + ax1];
        area += (ax * by - ay * bx) * 0.5f;
    }
    std::vector<int> rest(idx, idx + n);
    size_t start = 0;
    int trips = 10;
    while (rest.size() > 3 && trips > 0) {
        const size_t m = rest.size();
        if (start >= m) { --trips; start -= m; }
        int ear[3];
        float x[3], y[3];
        for (size_t k = 0; k < 3; ++k) {
            ear[k] = rest[(start + k) % m];
            if (ok(ear[k], ax0) && ok(ear[k], ax1)) {
                x[k] = pos[size_t(ear[k]) * 3 + ax0];
                y[k] = pos[size_t(ear[k]) * 3 + ax1];
            } else {
                x[k] = 0.f; y[k] = 0.f;
            }
        }
        const float e0x = x[1] - x[0], e0y = y[1] - y[0], e1x = x[2] - x[1], e1y = y[2] - y[1];
        const float cross = e0x * e1y - e0y * e1x;
        if (cross * area < 0.f) { ++start; continue; }  // reflex corner
        bool blocked = false;
        for (size_t o = 3; o < m && !blocked; ++o) {
            const int vi = rest[(start + o) % m];
            if (!ok(vi, ax0) || !ok(vi, ax1)) continue;
            blocked = inside_triangle(x, y, pos[size_t(vi) * 3 + ax0], pos[size_t(vi) * 3 + ax1]);
        }
        if (blocked) { ++start; continue; }
        tri.insert(tri.end(), ear, ear + 3);
        rest.erase(rest.begin() + static_cast<long>((start + 1) % m));
    }
    if (rest.size() == 3) tri.insert(tri.end(), rest.begin(), rest.end());
}

// triangulates the pending faces against the first `visible_vertices` vertices of the file
void flush_faces(ObjReader& r, size_t visible_vertices) {
    uint32_t begin = 0;
    for (uint32_t end : r.face_end) {
        split_face(r.corners.data() + begin, end - begin, r.pos.data(), visible_vertices * 3, r.tri);
        begin = end;
    }
    r.corners.clear();
    r.face_end.clear();
}

// the statements of one piece of the text (runs on its own thread)
void parse_piece(ObjPiece& pc) {
    const size_t len = static_cast<size_t>(pc.end - pc.begin);
    pc.pos.reserve(len / 24);
    pc.corners.reserve(len / 16);
    pc.face_end.reserve(len / 48);
    const char* p = pc.begin;
    const char* const eof = pc.end;
    while (p < eof) {
        // a line ends at LF, CR or CR LF; a NUL byte hides the rest of its line
        const char* nl = p;
        while (nl < eof && *nl != '\n' && *nl != '\r') ++nl;
        const char* next = nl;
        if (next < eof) next += (*next == '\r' && next + 1 < eof && next[1] == '\n') ? 2 : 1;
        const char* le = static_cast<const char*>(memchr(p, 0, static_cast<size_t>(nl - p)));
        if (!le) le = nl;
        ++pc.lines;
        const char* t = p;
        p = next;
        while (t < le && is_blank(*t)) ++t;
        if (t >= le || *t == '#') continue;
        const char c0 = t[0];
        const char c1 = t + 1 < le ? t[1] : '\0';
        if (c0 == 'v' && is_blank(c1)) {
            t += 2;
            const float x = next_real(t, le), y = next_real(t, le), z = next_real(t, le);
            pc.pos.push_back(x); pc.pos.push_back(y); pc.pos.push_back(z);
        } else if (c0 == 'f' && is_blank(c1)) {
            t += 2;
            while (t < le && is_blank(*t)) ++t;
            const long long nv = static_cast<long long>(pc.pos.size() / 3);
            while (t < le) {
                int v;
                if (!read_corner(t, le, &v)) {
                    pc.bad_line = pc.lines;  // tinyobjloader refuses the file here: nothing after this line matters
                    return;
                }
                pc.corners.push_back(v > 0 ? static_cast<long long>(v) - 1 : kRelative + nv + v);
                while (t < le && is_blank(*t)) ++t;
            }
            pc.face_end.push_back(static_cast<uint32_t>(pc.corners.size()));
        } else if ((c0 == 'g' || c0 == 'o') && is_blank(c1)) {
            pc.flushes.push_back({static_cast<uint32_t>(pc.face_end.size()), static_cast<uint32_t>(pc.pos.size() / 3)});
        }
        // `usemtl` is not a flush point here.  tinyobjloader flushes when the material ID changes, and IDs come from
        // an MTL file it resolves against the working directory (Mesh::load passes no base directory, mesh.cpp:192):
        // the reference's models ship without MTL files, every name maps to "no material" and nothing is flushed.
        // Only a polygon that names vertices defined after it could tell the difference.
        // vn, vt, s, t, mtllib and unknown statements do not reach Mesh::load's output
    }
}

// pieces of whole lines, about equal in bytes; a CR LF pair is never split
std::vector<ObjPiece> cut_text(const char* text, size_t len, size_t pieces) {
    std::vector<ObjPiece> out(pieces ? pieces : 1);
    const char* const eof = text + len;
    const char* p = text;
    for (size_t k = 0; k < out.size(); ++k) {
        out[k].begin = p;
        const char* q = (k + 1 == out.size()) ? eof : text + len / out.size() * (k + 1);
        if (q < p) q = p;
        while (q < eof && q > text && q[-1] != '\n' && q[-1] != '\r') ++q;
        if (q < eof && q > text && q[-1] == '\r' && *q == '\n') ++q;
        if (q == text && k + 1 != out.size()) q = p;
        out[k].end = p = q;
    }
    return out;
}

size_t obj_threads(size_t len) {
    if (const char* env = getenv("RTR_OBJ_THREADS")) {  // tests cut small files into many pieces with this
        const long n = atol(env);
        if (n >= 1) return static_cast<size_t>(n > 256 ? 256 : n);
    }
    size_t n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    if (n > 32) n = 32;
    const size_t by_size = len / (1u << 20);  // a piece is worth a thread from about 1 MB
    return by_size < 1 ? 1 : (by_size < n ? by_size : n);
}

// fn(k) for k in [0, n) on n threads (the caller's included)
template <class F> void run_on_threads(size_t n, F fn) {
    std::vector<std::thread> workers;
    for (size_t k = 1; k < n; ++k) workers.emplace_back(fn, k);
    fn(static_cast<size_t>(0));
    for (std::thread& w : workers) w.join();
}

int parse_obj(const char* text, size_t len, uint32_t model_id, rtr_triangle** out_tris, uint64_t* out_n) {
    // 1. the text, in parallel: numbers and corner lists of every piece
    std::vector<ObjPiece> pieces = cut_text(text, len, obj_threads(len));
    run_on_threads(pieces.size(), [&](size_t k) { parse_piece(pieces[k]); });
    uint64_t lines_before = 0;
    for (const ObjPiece& pc : pieces) {  // the first malformed `f` of the file, in file order
        if (pc.bad_line)
            return rtr_set_error(nullptr, RTR_E_INVALID, "obj: line %llu: bad `f` corner (index 0 or not a number)",
                                 static_cast<unsigned long long>(lines_before + pc.bad_line));
        lines_before += pc.lines;
    }
    // 2. in file order: vertex numbering, relative indices, flush points, polygons
    ObjReader r;
    size_t total_floats = 0, total_corners = 0;
    for (const ObjPiece& pc : pieces) { total_floats += pc.pos.size(); total_corners += pc.corners.size(); }
    r.pos.resize(total_floats);
    r.tri.reserve(total_corners);
    {
        std::vector<size_t> at(pieces.size() + 1, 0);
        for (size_t k = 0; k < pieces.size(); ++k) at[k + 1] = at[k] + pieces[k].pos.size();
        run_on_threads(pieces.size(), [&](size_t k) {
            if (!pieces[k].pos.empty()) memcpy(r.pos.data() + at[k], pieces[k].pos.data(), pieces[k].pos.size() * sizeof(float));
        });
    }
    size_t vertices_before = 0;
    for (const ObjPiece& pc : pieces) {
        size_t next_flush = 0;
        uint32_t begin = 0;
        for (uint32_t f = 0; f <= pc.face_end.size(); ++f) {
            while (next_flush < pc.flushes.size() && pc.flushes[next_flush].faces_before == f)
                flush_faces(r, vertices_before + pc.flushes[next_flush++].vertices_before);
            if (f == pc.face_end.size()) break;
            const uint32_t end = pc.face_end[f];
            if (end - begin == 3 && r.face_end.empty()) {  // a triangle with nothing pending before it passes straight through
                for (uint32_t c = begin; c < end; ++c) {
                    const long long v = pc.corners[c];
                    r.tri.push_back(static_cast<int>(v >= kRelative / 2 ? v - kRelative + static_cast<long long>(vertices_before) : v));
                }
            } else {
                for (uint32_t c = begin; c < end; ++c) {
                    const long long v = pc.corners[c];
                    r.corners.push_back(static_cast<int>(v >= kRelative / 2 ? v - kRelative + static_cast<long long>(vertices_before) : v));
                }
                r.face_end.push_back(static_cast<uint32_t>(r.corners.size()));
            }
            begin = end;
        }
        vertices_before += pc.pos.size() / 3;
    }
    flush_faces(r, vertices_before);

    // 3. the records, in parallel again: every corner must name a vertex of the file (the reference reads out of bounds)
    const size_t nv = r.pos.size() / 3;
    const size_t nt = r.tri.size() / 3;
    rtr_triangle* tris = static_cast<rtr_triangle*>(malloc((nt ? nt : 1) * sizeof(rtr_triangle)));
    if (!tris) return rtr_set_error(nullptr, RTR_E_NOMEM, "obj: %zu triangles do not fit in host memory", nt);
    const size_t parts = pieces.size();
    std::vector<size_t> first_bad(parts, static_cast<size_t>(-1));
    run_on_threads(parts, [&](size_t k) {
        const size_t t0 = nt * k / parts, t1 = nt * (k + 1) / parts;
        for (size_t i = t0; i < t1; ++i) {
            rtr_triangle& out = tris[i];
            memset(&out, 0, sizeof(out));
            float* dst[3] = {out.p0, out.p1, out.p2};
            for (int c = 0; c < 3; ++c) {
                const int v = r.tri[3 * i + c];
                if (v < 0 || static_cast<size_t>(v) >= nv) {
                    if (first_bad[k] == static_cast<size_t>(-1)) first_bad[k] = 3 * i + c;
                    continue;
                }
                const float* src = &r.pos[static_cast<size_t>(v) * 3];
                dst[c][0] = src[0]; dst[c][1] = src[1]; dst[c][2] = src[2]; dst[c][3] = 1.f;
            }
            out.model_id = model_id;
        }
    });
    for (size_t k = 0; k < parts; ++k)
        if (first_bad[k] != static_cast<size_t>(-1)) {
            const size_t i = first_bad[k];
            free(tris);
            return rtr_set_error(nullptr, RTR_E_INVALID, "obj: triangle %zu names vertex %d of %zu", i / 3,
                                 r.tri[i] < 0 ? r.tri[i] : r.tri[i] + 1, nv);
        }
    *out_tris = tris;
    *out_n = nt;
    return RTR_OK;
}

// glm 0.9.9.9 mat4 * mat4 (type_mat4x4.inl:630-648), column-major: every entry is ((a0*b0 + a1*b1) + a2*b2) + a3*b3
void mul4(const float a[16], const float b[16], float out[16]) {
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r[4 * c + row] = ((a[row] * b[4 * c] + a[4 + row] * b[4 * c + 1]) + a[8 + row] * b[4 * c + 2]) + a[12 + row] * b[4 * c + 3];
    memcpy(out, r, sizeof(r));
}
void transpose4(const float a[16], float out[16]) {
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) r[4 * c + row] = a[4 * row + c];
    memcpy(out, r, sizeof(r));
}

}  // namespace

extern "C" {

int rtr_obj_parse(const char* text, uint64_t len, uint32_t model_id, rtr_triangle** out_tris, uint64_t* out_n) {
    if (!out_tris || !out_n) return rtr_set_error(nullptr, RTR_E_INVALID, "obj_parse: NULL output");
    if (!text && len) return rtr_set_error(nullptr, RTR_E_INVALID, "obj_parse: NULL text");
    *out_tris = nullptr;
    *out_n = 0;
    return parse_obj(text ? text : "", static_cast<size_t>(len), model_id, out_tris, out_n);
}

int rtr_obj_load(const char* path, uint32_t model_id, rtr_triangle** out_tris, uint64_t* out_n) {
    if (!out_tris || !out_n) return rtr_set_error(nullptr, RTR_E_INVALID, "obj_load: NULL output");
    *out_tris = nullptr;
    *out_n = 0;
    if (!path) return rtr_set_error(nullptr, RTR_E_INVALID, "obj_load: NULL path");
    FILE* f = fopen(path, "rb");
    if (!f) return rtr_set_error(nullptr, RTR_E_INVALID, "obj_load: cannot open file [%s]: %s", path, strerror(errno));
    std::vector<char> buf;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        if (sz > 0) buf.reserve(static_cast<size_t>(sz));
        rewind(f);
    }
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (bad) return rtr_set_error(nullptr, RTR_E_INVALID, "obj_load: read error on [%s]", path);
    return parse_obj(buf.data(), buf.size(), model_id, out_tris, out_n);
}

void rtr_obj_free(rtr_triangle* tris) { free(tris); }

int rtr_mesh_primitive(int which, uint32_t model_id, rtr_triangle* out, uint64_t cap, uint64_t* out_n) {
    // corner code: bit 0 = +x, bit 1 = +y, bit 2 = +z (else -1); the triangle and the square lie in z = 0
    static const uint8_t cube[36] = {2, 0, 1, 2, 1, 3,  4, 6, 5, 5, 6, 7,  5, 7, 3, 5, 3, 1,
                                     6, 4, 2, 2, 4, 0,  2, 3, 7, 2, 7, 6,  0, 5, 1, 0, 4, 5};
    static const uint8_t square[6] = {2, 0, 1, 2, 1, 3};
    if (!out_n) return rtr_set_error(nullptr, RTR_E_INVALID, "mesh_primitive: NULL out_n");
    const uint8_t* codes = nullptr;
    uint64_t n = 0;
    switch (which) {
        case RTR_PRIMITIVE_TRIANGLE: n = 1; break;
        case RTR_PRIMITIVE_SQUARE: n = 2; codes = square; break;
        case RTR_PRIMITIVE_CUBE: n = 12; codes = cube; break;
        case RTR_PRIMITIVE_SPHERE: n = 0; break;  // mesh.cpp:181-184 returns an empty mesh
        default: return rtr_set_error(nullptr, RTR_E_INVALID, "mesh_primitive: unknown primitive %d", which);
    }
    *out_n = n;
    if (n > cap || (n && !out)) return rtr_set_error(nullptr, RTR_E_INVALID, "mesh_primitive: %llu triangles, room for %llu",
                                                     (unsigned long long)n, (unsigned long long)cap);
    if (n) memset(out, 0, n * sizeof(rtr_triangle));
    for (uint64_t i = 0; i < n; ++i) {
        float* dst[3] = {out[i].p0, out[i].p1, out[i].p2};
        for (int k = 0; k < 3; ++k) {
            if (codes) {
                const uint8_t c = codes[3 * i + k];
                dst[k][0] = (c & 1) ? 1.f : -1.f;
                dst[k][1] = (c & 2) ? 1.f : -1.f;
                dst[k][2] = which == RTR_PRIMITIVE_CUBE ? ((c & 4) ? 1.f : -1.f) : 0.f;
            } else {  // (0,1,0), (-1,-1,0), (1,-1,0)
                dst[k][0] = k == 0 ? 0.f : k == 1 ? -1.f : 1.f;
                dst[k][1] = k == 0 ? 1.f : -1.f;
                dst[k][2] = 0.f;
            }
            dst[k][3] = 1.f;
        }
        out[i].model_id = model_id;
    }
    return RTR_OK;
}

void rtr_mesh_init(rtr_mesh* m) {
    if (!m) return;
    memset(m, 0, sizeof(*m));
    m->model[0] = m->model[5] = m->model[10] = m->model[15] = 1.f;
}
void rtr_mesh_set_model(rtr_mesh* m, const float model[16]) {
    if (m && model) memcpy(m->model, model, 64);
}
void rtr_mesh_set_material(rtr_mesh* m, uint32_t material_id) {
    if (m) m->material_id = material_id;
}
void rtr_mesh_set_position(rtr_mesh* m, float x, float y, float z) {
    if (!m) return;
    m->model[12] = x; m->model[13] = y; m->model[14] = z;
}
// M <- transpose(S * transpose(M)), S = diag(s, s, s, 1)   (mesh.cpp:29-35; a full matrix product, zeros included)
void rtr_mesh_set_scale(rtr_mesh* m, float scale) {
    if (!m) return;
    float s[16] = {0}, t[16];
    s[0] = s[5] = s[10] = 1.f * scale;
    s[15] = 1.f;
    transpose4(m->model, t);
    mul4(s, t, t);
    transpose4(t, m->model);
}
// M <- transpose(Rz * Ry * Rx * transpose(M)) with the matrices exactly as mesh.cpp:37-58 writes them: each braced
// row of the source is a COLUMN for glm, and cos/sin are the double functions rounded to float.
void rtr_mesh_set_rotation(rtr_mesh* m, float theta_x, float theta_y, float theta_z) {
    if (!m) return;
    const float cx = static_cast<float>(cos(static_cast<double>(theta_x))), sx = static_cast<float>(sin(static_cast<double>(theta_x)));
    const float cy = static_cast<float>(cos(static_cast<double>(theta_y))), sy = static_cast<float>(sin(static_cast<double>(theta_y)));
    const float cz = static_cast<float>(cos(static_cast<double>(theta_z))), sz = static_cast<float>(sin(static_cast<double>(theta_z)));
    const float rx[16] = {1.f, 0.f, 0.f, 0.f,  0.f, cx, -sx, 0.f,  0.f, sx, cx, 0.f,  0.f, 0.f, 0.f, 1.f};
    const float ry[16] = {cy, 0.f, sy, 0.f,  0.f, 1.f, 0.f, 0.f,  -sy, 0.f, cy, 0.f,  0.f, 0.f, 0.f, 1.f};
    const float rz[16] = {cz, -sz, 0.f, 0.f,  sz, cz, 0.f, 0.f,  0.f, 0.f, 1.f, 0.f,  0.f, 0.f, 0.f, 1.f};
    float t[16], acc[16];
    mul4(rz, ry, acc);
    mul4(acc, rx, acc);
    transpose4(m->model, t);
    mul4(acc, t, t);
    transpose4(t, m->model);
}

// vec3((1/3) * model * (P0 + P1 + P2))  (triangle.cpp:30-32): scalar * mat4 first, then the matrix-vector product
void rtr_triangle_centroid(const rtr_triangle* t, const float model[16], float out[3]) {
    if (!t || !model || !out) return;
    const float third = 1.f / 3.f;
    float sm[16], s[4];
    for (int i = 0; i < 16; ++i) sm[i] = model[i] * third;
    for (int k = 0; k < 4; ++k) s[k] = (t->p0[k] + t->p1[k]) + t->p2[k];
    for (int row = 0; row < 3; ++row)
        out[row] = ((sm[row] * s[0] + sm[4 + row] * s[1]) + (sm[8 + row] * s[2] + sm[12 + row] * s[3]));
}

// cr::Camera for callers that cannot include rtr_scene.hpp (the Python mirror): a camera made like
// application.cpp:16-20, the input events replayed on it -- kind 0: ProcessMouseMovement(a, b); kind 1..6:
// processKeyboard(FORWARD..DOWN, a) with _Accelerate = (b != 0) -- then getGpuData() (camera.cpp:23-34).
int rtr_camera_gpu_data(const float eye[3], float aspect, float fov, float near_plane, float far_plane, int n_events,
                        const int* kind, const float* a, const float* b, rtr_camera* out) {
    if (!eye || !out || n_events < 0 || (n_events && (!kind || !a || !b)))
        return rtr_set_error(nullptr, RTR_E_INVALID, "camera_gpu_data: NULL argument");
    cr::Camera cam(eye, aspect, fov, near_plane, far_plane);
    for (int i = 0; i < n_events; ++i) {
        if (kind[i] == 0) cam.ProcessMouseMovement(a[i], b[i]);
        else if (kind[i] >= 1 && kind[i] <= 6) {
            cam._Accelerate = b[i] != 0.f;
            cam.processKeyboard(static_cast<cr::CameraMovement>(kind[i] - 1), a[i]);
        } else return rtr_set_error(nullptr, RTR_E_INVALID, "camera_gpu_data: event %d has kind %d", i, kind[i]);
    }
    const cr::CameraGPU g = cam.getGpuData();
    memcpy(out, &g, sizeof(g));
    return RTR_OK;
}

}  // extern "C"
