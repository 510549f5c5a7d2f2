// mesh_host.cpp — TGRID reader, geometry and per-mesh precomputation (host side of liborc_b200).
// Behaviour follows src/io.rs:32-515 of the reference (the subset of TGRID it accepts is listed in
// SURVEY.md Appendix A); data layout and algorithms are this library's own: flat SoA arrays and a
// single pass over the file buffer instead of per-entity hash maps.
#include "mesh_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

#include "../../include/orc_b200.h"
#include "vecmath.cuh"

namespace orc {

// Host threads for the per-face / per-cell loops below (every item writes only its own slots; sums keep their sequential order
// inside an item, so the results do not depend on the thread count). ORC_B200_HOST_THREADS overrides the default (<= 16).
static int host_threads() {
    static const int n = [] {
        if (const char* e = getenv("ORC_B200_HOST_THREADS")) return std::max(1, atoi(e));
        unsigned h = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(h ? h : 1u, 16u));
    }();
    return n;
}
template <class F>
static void parallel_for(int64_t n, F body) {  // body(begin, end) over disjoint ranges; the first exception is rethrown
    const int64_t grain = 16384;
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), (n + grain - 1) / grain));
    if (T == 1) { body((int64_t)0, n); return; }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    for (int t = 0; t < T; ++t) {
        const int64_t b = n * t / T, e = n * (t + 1) / T;
        th.emplace_back([&, b, e] {
            try { body(b, e); } catch (...) { std::lock_guard<std::mutex> g(mu); if (!err) err = std::current_exception(); }
        });
    }
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

int HostMesh::find_zone(const std::string& name) const {
    for (size_t k = 0; k < zones.size(); ++k)
        if (zones[k].name == name) return (int)k;
    return -1;
}

static bool known_bc_id(int64_t t) {  // src/mesh.rs:50-66
    static const int64_t ids[] = {2, 3, 4, 5, 7, 8, 9, 10, 12, 14, 20, 24, 31, 36, 37};
    for (int64_t v : ids)
        if (v == t) return true;
    return false;
}

// ---- geometry: src/io.rs:289-438 ----------------------------------------------------------------
// Input: face_nodes / face_node_ptr, raw face_c0 / face_c1 (-1 = "cell 0" of the file), zones.
static void build_geometry(HostMesh& m) {
    const int dims = m.dims;
    const int64_t F = m.n_faces;
    if (dims != 2 && dims != 3) throw MeshError(ORC_E_IO, "dimensions must be 2 or 3");
    int64_t max_cell = -1;
    for (int64_t f = 0; f < F; ++f) max_cell = std::max<int64_t>(max_cell, std::max(m.face_c0[f], m.face_c1[f]));
    m.n_cells = max_cell + 1;
    const int64_t N = m.n_cells;
    m.face_area.assign(F, 0.);
    m.face_flipped.assign(F, 0);
    m.face_normal.assign(3 * F, 0.);
    m.face_centroid.assign(3 * F, 0.);
    m.cell_volume.assign(N, 0.);
    m.cell_centroid.assign(3 * N, 0.);
    auto P = [&](int64_t f, int k) -> V3 {
        int64_t n = m.face_nodes[m.face_node_ptr[f] + k];
        return v3(m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]);
    };
    // faces: every face writes its own slots
    parallel_for(F, [&](int64_t f0, int64_t f1) {
        for (int64_t f = f0; f < f1; ++f) {
            const int cnt = (int)(m.face_node_ptr[f + 1] - m.face_node_ptr[f]);
            if (cnt < dims) throw MeshError(ORC_E_IO, "face has too few nodes");
            for (int k = 0; k < cnt; ++k) {
                int64_t n = m.face_nodes[m.face_node_ptr[f] + k];
                if (n < 0 || n >= m.n_nodes) throw MeshError(ORC_E_IO, "nodes should have all been read");
            }
            V3 nrm;
            if (dims == 2) {  // io.rs:305-321
                V3 t = vsub(P(f, 1), P(f, 0));
                nrm = (t.x == 0.) ? v3(1., -t.x / t.y, 0.) : v3(-t.y / t.x, 1., 0.);
                nrm = vunit(nrm);
            } else {  // io.rs:322-326
                nrm = vunit(vcross(vsub(P(f, 2), P(f, 1)), vsub(P(f, 1), P(f, 0))));
            }
            if (m.face_c0[f] < 0) {  // io.rs:332-337: no cell 0 -> flip the normal, keep the other cell as cell_indices[0]
                nrm = vneg(nrm);
                m.face_c0[f] = m.face_c1[f];
                m.face_c1[f] = -1;
                m.face_flipped[f] = 1;
            }
            V3 acc = vzero();  // io.rs:338-342
            for (int k = 0; k < cnt; ++k) acc = vadd(acc, P(f, k));
            V3 cen = vdivs(acc, (double)cnt);
            double area;
            if (cnt == 2) {  // io.rs:345-349
                if (dims != 2) throw MeshError(ORC_E_IO, "assertion failed: dimensions == 2");
                area = vnorm(vsub(P(f, 1), P(f, 0)));
            } else {  // io.rs:375-396: triangle fan about the centroid
                auto tri = [](V3 a, V3 b, V3 c) { return fabs(vnorm(vcross(vsub(b, a), vsub(c, a)))) / 2.; };
                area = 0.;
                for (int k = 0; k + 1 < cnt; ++k) area = area + tri(cen, P(f, k), P(f, k + 1));
                area = area + tri(cen, P(f, 0), P(f, cnt - 1));
            }
            m.face_area[f] = area;
            m.face_normal[3 * f] = nrm.x; m.face_normal[3 * f + 1] = nrm.y; m.face_normal[3 * f + 2] = nrm.z;
            m.face_centroid[3 * f] = cen.x; m.face_centroid[3 * f + 1] = cen.y; m.face_centroid[3 * f + 2] = cen.z;
        }
    });
    // cell -> faces in ascending face index (the order of the pushes at io.rs:410)
    std::vector<int32_t> nfaces(N, 0);
    for (int64_t f = 0; f < F; ++f) {
        if (m.face_c0[f] >= 0) nfaces[m.face_c0[f]]++;
        if (m.face_c1[f] >= 0) nfaces[m.face_c1[f]]++;
    }
    m.cf_ptr.assign(N + 1, 0);
    for (int64_t c = 0; c < N; ++c) {
        if (nfaces[c] == 0) throw MeshError(ORC_E_IO, "cell index missing from mesh");
        m.cf_ptr[c + 1] = m.cf_ptr[c] + nfaces[c];
    }
    m.cf_face.assign(m.cf_ptr[N], 0);
    {
        std::vector<int32_t> pos(m.cf_ptr.begin(), m.cf_ptr.end() - 1);
        for (int64_t f = 0; f < F; ++f) {
            if (m.face_c0[f] >= 0) m.cf_face[pos[m.face_c0[f]]++] = (int32_t)f;
            if (m.face_c1[f] >= 0) m.cf_face[pos[m.face_c1[f]]++] = (int32_t)f;
        }
    }
    // cells: the centroid is the mean of the face centroids, summed in ascending face index like the pushes of io.rs:404-414
    parallel_for(N, [&](int64_t c0, int64_t c1) {
        for (int64_t c = c0; c < c1; ++c) {  // io.rs:417-438
            V3 sum = vzero();
            for (int32_t q = m.cf_ptr[c]; q < m.cf_ptr[c + 1]; ++q) {
                int64_t f = m.cf_face[q];
                sum = vadd(sum, v3(m.face_centroid[3 * f], m.face_centroid[3 * f + 1], m.face_centroid[3 * f + 2]));
            }
            V3 cc = vdivs(sum, (double)nfaces[c]);
            if (nfaces[c] < dims + 1) throw MeshError(ORC_E_IO, "cell has too few faces");
            double vol = 0.;
            for (int32_t q = m.cf_ptr[c]; q < m.cf_ptr[c + 1]; ++q) {
                int64_t f = m.cf_face[q];
                V3 fc = v3(m.face_centroid[3 * f], m.face_centroid[3 * f + 1], m.face_centroid[3 * f + 2]);
                V3 fn = v3(m.face_normal[3 * f], m.face_normal[3 * f + 1], m.face_normal[3 * f + 2]);
                vol = vol + m.face_area[f] * fabs(vdot(vsub(fc, cc), fn)) / (double)dims;
            }
            m.cell_volume[c] = vol;
            m.cell_centroid[3 * c] = cc.x; m.cell_centroid[3 * c + 1] = cc.y; m.cell_centroid[3 * c + 2] = cc.z;
        }
    });
}

// ---- per-mesh precomputation for the device path -------------------------------------------------
static void build_derived(HostMesh& m) {
    const int64_t N = m.n_cells;
    const int64_t S = m.cf_ptr[N];
    m.cf_nb.assign(S, -1);
    m.cf_slot.assign(S, -1);
    m.rowptr.assign(N + 1, 0);
    // neighbour per (cell, face-slot): the other cell of a two-cell face
    parallel_for(N, [&](int64_t c0, int64_t c1) {
        for (int64_t c = c0; c < c1; ++c)
            for (int32_t q = m.cf_ptr[c]; q < m.cf_ptr[c + 1]; ++q) {
                int32_t f = m.cf_face[q];
                if (m.face_c1[f] >= 0) m.cf_nb[q] = (m.face_c0[f] == (int32_t)c) ? m.face_c1[f] : m.face_c0[f];
            }
    });
    // pattern: {i} U neighbours, sorted, duplicates merged (CsrMatrix::from(&CooMatrix) sums duplicates); two passes: row
    // lengths, then the columns
    const int64_t olo = m.own_hi < 0 ? 0 : m.own_lo, ohi = m.own_hi < 0 ? N : m.own_hi;
    auto row_of = [&](int64_t c, std::vector<int32_t>& tmp) {
        tmp.clear();
        if (c < olo || c >= ohi) return;  // halo cell: no matrix row
        tmp.push_back((int32_t)c);
        for (int32_t q = m.cf_ptr[c]; q < m.cf_ptr[c + 1]; ++q)
            if (m.cf_nb[q] >= 0) tmp.push_back(m.cf_nb[q]);
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    };
    std::vector<int32_t> rowlen(N, 0);
    parallel_for(N, [&](int64_t c0, int64_t c1) {
        std::vector<int32_t> tmp;
        for (int64_t c = c0; c < c1; ++c) { row_of(c, tmp); rowlen[c] = (int32_t)tmp.size(); }
    });
    int64_t total = 0;
    for (int64_t c = 0; c < N; ++c) {
        total += rowlen[c];
        if (total > INT32_MAX) throw MeshError(ORC_E_INVALID, "pattern exceeds 2^31 entries");
        m.rowptr[c + 1] = (int32_t)total;
    }
    m.col.assign((size_t)total, 0);
    m.diag_idx.assign(N, -1);
    parallel_for(N, [&](int64_t c0, int64_t c1) {
        std::vector<int32_t> tmp;
        for (int64_t c = c0; c < c1; ++c) {
            row_of(c, tmp);
            std::copy(tmp.begin(), tmp.end(), m.col.begin() + m.rowptr[c]);
            const int32_t* b = m.col.data() + m.rowptr[c];
            const int32_t* e = m.col.data() + m.rowptr[c + 1];
            if (b == e) continue;  // halo cell
            m.diag_idx[c] = (int32_t)(std::lower_bound(b, e, (int32_t)c) - m.col.data());
            for (int32_t q = m.cf_ptr[c]; q < m.cf_ptr[c + 1]; ++q)
                if (m.cf_nb[q] >= 0) m.cf_slot[q] = (int32_t)(std::lower_bound(b, e, m.cf_nb[q]) - m.col.data());
        }
    });
    // level schedule of the in-place diagonal recurrence: cell i reads the NEW diagonal of every
    // neighbour j < i (src/discretization.rs:184-197 -> src/solver.rs:1068-1081, written at :340-351),
    // so level(i) = 1 + max level(j), j < i. Cells of one level are independent.
    m.level_of_cell.assign(N, 0);
    int32_t nlev = N > 0 ? 1 : 0;
    for (int64_t c = olo; c < ohi; ++c) {
        int32_t l = 0;
        for (int32_t k = m.rowptr[c]; k < m.diag_idx[c]; ++k)
            if (m.col[k] >= olo) l = std::max(l, m.level_of_cell[m.col[k]] + 1);  // lower-HALO neighbours are partition-lagged (C4)
        m.level_of_cell[c] = l;
        nlev = std::max(nlev, l + 1);
    }
    m.level_ptr.assign(nlev + 1, 0);
    for (int64_t c = olo; c < ohi; ++c) m.level_ptr[m.level_of_cell[c] + 1]++;
    for (int32_t l = 0; l < nlev; ++l) m.level_ptr[l + 1] += m.level_ptr[l];
    m.level_order.assign(ohi - olo, 0);
    {
        std::vector<int32_t> pos(m.level_ptr.begin(), m.level_ptr.end() - 1);
        for (int64_t c = olo; c < ohi; ++c) m.level_order[pos[m.level_of_cell[c]]++] = (int32_t)c;  // ascending cell id inside a level
    }
}

// ---- TGRID ASCII reader: src/io.rs:32-287 --------------------------------------------------------
namespace {
struct Lines {  // line cursor over a file buffer; '\r' before '\n' is dropped
    const char* p;
    const char* end;
    bool next(const char*& b, const char*& e) {
        if (p >= end) return false;
        b = p;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        e = nl ? nl : end;
        p = nl ? nl + 1 : end;
        if (e > b && e[-1] == '\r') --e;
        return true;
    }
};
inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f'; }
// split_ascii_whitespace
inline int split(const char* b, const char* e, const char** tb, const char** te, int cap) {
    int n = 0;
    while (b < e) {
        while (b < e && is_ws(*b)) ++b;
        const char* s = b;
        while (b < e && !is_ws(*b)) ++b;
        if (b > s) {
            if (n < cap) { tb[n] = s; te[n] = b; }
            ++n;
        }
    }
    return n;
}
// usize::from_str_radix(s, 16): hex digits of either case, optional leading '+'
inline bool hex_token(const char* b, const char* e, uint64_t& out) {
    if (b < e && *b == '+') ++b;
    if (b >= e) return false;
    uint64_t v = 0;
    for (; b < e; ++b) {
        int d;
        if (*b >= '0' && *b <= '9') d = *b - '0';
        else if (*b >= 'a' && *b <= 'f') d = *b - 'a' + 10;
        else if (*b >= 'A' && *b <= 'F') d = *b - 'A' + 10;
        else return false;
        v = v * 16 + (uint64_t)d;
    }
    out = v;
    return true;
}
// the header regex ([0-9a-z]+) over the whole line, every capture parsed as hex (io.rs:47-54)
std::vector<uint64_t> header_items(const char* b, const char* e) {
    std::vector<uint64_t> items;
    auto ok = [](char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z'); };
    while (b < e) {
        while (b < e && !ok(*b)) ++b;
        const char* s = b;
        while (b < e && ok(*b)) ++b;
        if (b > s) {
            uint64_t v;
            if (!hex_token(s, b, v)) throw MeshError(ORC_E_IO, "valid hex");
            items.push_back(v);
        }
    }
    return items;
}
}  // namespace

// Fast path for the body of a face section whose faces all have `node_count` nodes (the (13 header's face type 2 / 3 / 4: every
// section of the meshes this solver is fed): the body [body, the first line that starts with ')') is cut at line boundaries and
// parsed by the host threads into `out_c` (2 per face) and `out_nodes` (node_count per face). Returns false WITHOUT side effects
// if anything is irregular (a "(" line after the first, a wrong token count, a bad hex token, no terminator): the caller then
// runs the sequential parser, which reproduces the reference's behaviour on such input line by line (io.rs:194-274).
static bool parse_uniform_faces(const char* body, const char* file_end, int node_count, std::vector<int32_t>& out_c,
                                std::vector<int32_t>& out_nodes, const char*& after_section) {
    const char* p = body;
    {   // an opening "(" line of its own is skipped, like the sequential parser does
        const char* nl = (const char*)memchr(p, '\n', (size_t)(file_end - p));
        if (!nl) return false;
        const char* e = (nl > p && nl[-1] == '\r') ? nl - 1 : nl;
        if (e - p == 1 && *p == '(') p = nl + 1;
    }
    if (p >= file_end) return false;
    const char* sec_end;   // start of the terminating line
    if (*p == ')') sec_end = p;
    else {
        const char* hit = (const char*)memmem(p, (size_t)(file_end - p), "\n)", 2);
        if (!hit) return false;
        sec_end = hit + 1;
    }
    const char* term_nl = (const char*)memchr(sec_end, '\n', (size_t)(file_end - sec_end));
    after_section = term_nl ? term_nl + 1 : file_end;
    const int64_t bytes = sec_end - p;
    if (bytes <= 0) { out_c.clear(); out_nodes.clear(); return true; }
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), bytes / (1 << 20) + 1));
    std::vector<const char*> cut(T + 1);
    cut[0] = p; cut[T] = sec_end;
    for (int t = 1; t < T; ++t) {
        const char* q = p + bytes * t / T;
        const char* nl = (const char*)memchr(q, '\n', (size_t)(sec_end - q));
        cut[t] = nl ? nl + 1 : sec_end;
    }
    for (int t = 1; t <= T; ++t) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<int64_t> lines(T + 1, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                int64_t n = 0;
                for (const char* q = cut[t]; q < cut[t + 1];) {
                    const char* nl = (const char*)memchr(q, '\n', (size_t)(cut[t + 1] - q));
                    ++n;
                    q = nl ? nl + 1 : cut[t + 1];
                }
                lines[t + 1] = n;
            });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; ++t) lines[t + 1] += lines[t];
    const int64_t total = lines[T];
    std::vector<int32_t> c((size_t)total * 2), nodes((size_t)total * node_count);
    std::vector<char> bad(T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                const char *tb[66], *te[66];
                int64_t idx = lines[t];
                for (const char* q = cut[t]; q < cut[t + 1]; ++idx) {
                    const char* nl = (const char*)memchr(q, '\n', (size_t)(cut[t + 1] - q));
                    const char* e = nl ? nl : cut[t + 1];
                    const char* next = nl ? nl + 1 : cut[t + 1];
                    if (e > q && e[-1] == '\r') --e;
                    const int n = split(q, e, tb, te, 66);
                    if (n != node_count + 2) { bad[t] = 1; return; }
                    for (int k = 0; k < 2; ++k) {
                        uint64_t v;
                        if (!hex_token(tb[node_count + k], te[node_count + k], v) || v > (uint64_t)INT32_MAX) { bad[t] = 1; return; }
                        c[2 * idx + k] = v > 0 ? (int32_t)(v - 1) : -1;
                    }
                    for (int k = 0; k < node_count; ++k) {
                        uint64_t v;
                        if (!hex_token(tb[k], te[k], v) || v > (uint64_t)INT32_MAX) { bad[t] = 1; return; }
                        nodes[idx * node_count + k] = v > 0 ? (int32_t)(v - 1) : -1;
                    }
                    q = next;
                }
            });
        for (auto& x : th) x.join();
    }
    for (char b : bad) if (b) return false;
    out_c.swap(c);
    out_nodes.swap(nodes);
    return true;
}

HostMesh* read_tgrid(const std::string& path) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw MeshError(ORC_E_IO, "Unable to open mesh file for reading.");
    std::vector<char> buf;
    {
        fseek(fp, 0, SEEK_END);
        long sz = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        buf.resize(sz > 0 ? (size_t)sz : 0);
        size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), fp);
        fclose(fp);
        buf.resize(got);
    }
    const bool debug = getenv("ORC_B200_DEBUG") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!debug) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[orc_b200] read_tgrid: %s %.3f s\n", what, std::chrono::duration<double>(t - t_start).count());
        t_start = t;
    };
    lap("file read");
    std::unique_ptr<HostMesh> mp(new HostMesh());
    HostMesh& m = *mp;
    Lines L{buf.data(), buf.data() + buf.size()};
    const char *hb, *he;
    if (!L.next(hb, he)) throw MeshError(ORC_E_IO, "mesh is at least one line long");
    int dims = 0;
    std::string zone_name;
    // entities arrive keyed by their (1-based) file index; stored sparsely then checked for gaps like the
    // reference's 0..len() lookups (io.rs:289-290, 494-506)
    std::vector<double> xyz;
    std::vector<char> have_node;
    struct RawFace { int64_t zone; int32_t c[2]; int64_t node_begin; int32_t node_count; };
    std::vector<RawFace> faces;
    std::vector<char> have_face;
    std::vector<int32_t> fnodes;
    const char *tb[64], *te[64];
    for (;;) {
        int nt = split(hb, he, tb, te, 64);
        if (nt == 0) throw MeshError(ORC_E_IO, "index out of bounds: empty section header line");  // section_header_blocks[0]
        std::string tag(tb[0], te[0]);
        auto zone_zero = [&]() {
            if (nt < 2) throw MeshError(ORC_E_IO, "index out of bounds: section header has one item");
            return (te[1] - tb[1] == 2) && tb[1][0] == '(' && tb[1][1] == '0';
        };
        if (tag == "(0") {  // io.rs:83-90
            const char* sp = he;
            while (sp > hb && sp[-1] != ' ') --sp;
            if (sp == hb) throw MeshError(ORC_E_IO, "comment has a space");
            zone_name.assign(sp, he);
            while (zone_name.size() >= 2 && zone_name.compare(zone_name.size() - 2, 2, "\")") == 0) zone_name.resize(zone_name.size() - 2);
        } else if (tag == "(2") {  // io.rs:92-104
            if (nt < 2 || te[1][-1] != ')') throw MeshError(ORC_E_IO, "dimensions section should have two items");
            dims = atoi(std::string(tb[1], te[1] - 1).c_str());
            if (dims != 2 && dims != 3) throw MeshError(ORC_E_IO, "Mesh is not 2D or 3D.");
        } else if (tag == "(10" && !zone_zero()) {  // io.rs:105-175
            std::vector<uint64_t> items = header_items(hb, he);
            if (items.size() != 6) throw MeshError(ORC_E_IO, "nodes header has six items");
            uint64_t node_number = items[2];
            const char *b, *e;
            if (!L.next(b, e)) throw MeshError(ORC_E_IO, "node section shouldn't be empty");
            for (;;) {
                if (e - b == 1 && *b == '(') {
                    if (!L.next(b, e)) throw MeshError(ORC_E_IO, "unexpected end of node section");
                    continue;
                }
                if (e > b && *b == ')') break;
                int n = split(b, e, tb, te, 64);
                if (n == dims) {
                    if (node_number == 0) throw MeshError(ORC_E_IO, "node index underflow");
                    uint64_t idx = node_number - 1;
                    if (idx >= have_node.size()) { have_node.resize(idx + 1, 0); xyz.resize(3 * (idx + 1), 0.); }
                    char* endp;
                    for (int k = 0; k < 3; ++k) {
                        double v = 0.;
                        if (k < dims) {
                            std::string s(tb[k], te[k]);
                            v = strtod(s.c_str(), &endp);
                            if (endp == s.c_str() || *endp) throw MeshError(ORC_E_IO, s + " should be a string representation of a float");
                        }
                        xyz[3 * idx + k] = v;
                    }
                    have_node[idx] = 1;
                }
                if (!L.next(b, e)) break;
                node_number += 1;
            }
            lap("node section");
        } else if (tag == "(12" && !zone_zero()) {  // io.rs:180-193 (cell zones are recorded but unused on the path)
            if (header_items(hb, he).size() != 6) throw MeshError(ORC_E_IO, "cell section has 6 entries");
        } else if (tag == "(13" && !zone_zero()) {  // io.rs:194-274
            std::vector<uint64_t> items = header_items(hb, he);
            if (items.size() != 6) throw MeshError(ORC_E_IO, "face section has 6 entries");
            const int64_t zone_id = (int64_t)items[1];
            const uint64_t start_index = items[2], bc = items[4], face_type = items[5];
            if (!known_bc_id((int64_t)bc)) throw MeshError(ORC_E_IO, "valid BC type");
            bool known = false;
            for (auto& z : m.zones) known = known || z.id == zone_id;
            if (!known) {  // entry().or_insert()
                HostZone z; z.id = zone_id; z.type = (int32_t)bc; z.name = zone_name;
                m.zones.push_back(z);
            }
            if (face_type >= 2 && face_type <= 4 && start_index > 0 && L.p < L.end) {  // uniform section: parsed by the host threads
                std::vector<int32_t> fc, fn;
                const char* after = nullptr;
                const bool fast = parse_uniform_faces(L.p, L.end, (int)face_type, fc, fn, after);
                if (debug) fprintf(stderr, "[orc_b200] read_tgrid: face section of zone %lld: %s\n", (long long)zone_id, fast ? "parallel fast path" : "sequential (irregular body)");
                if (fast) {
                    const uint64_t count = fc.size() / 2, first = start_index - 1;
                    if (first + count > have_face.size()) { have_face.resize(first + count, 0); faces.resize(first + count); }
                    const int64_t nb = (int64_t)fnodes.size();
                    fnodes.insert(fnodes.end(), fn.begin(), fn.end());
                    for (uint64_t k = 0; k < count; ++k) {
                        RawFace rf;
                        rf.zone = zone_id; rf.c[0] = fc[2 * k]; rf.c[1] = fc[2 * k + 1];
                        rf.node_begin = nb + (int64_t)k * (int64_t)face_type; rf.node_count = (int32_t)face_type;
                        faces[first + k] = rf;
                        have_face[first + k] = 1;
                    }
                    L.p = after;
                    if (!L.next(hb, he)) break;
                    continue;
                }
            }
            const char *b, *e;
            if (!L.next(b, e)) throw MeshError(ORC_E_IO, "face section has contents");
            uint64_t face_number = start_index;
            for (;;) {
                if (e - b == 1 && *b == '(') {
                    if (!L.next(b, e)) throw MeshError(ORC_E_IO, "unexpected end of face section");
                    continue;
                }
                if (e > b && *b == ')') break;
                int n = split(b, e, tb, te, 64);
                if (n > 64) throw MeshError(ORC_E_IO, "face line has more than 62 nodes");
                if (n < 2) break;
                const int node_count = n - 2;
                if (face_type != 0 && face_type != 5 && face_type != (uint64_t)node_count) break;
                if (face_number == 0) throw MeshError(ORC_E_IO, "face index underflow");
                uint64_t idx = face_number - 1;
                if (idx >= have_face.size()) { have_face.resize(idx + 1, 0); faces.resize(idx + 1); }
                RawFace rf;
                rf.zone = zone_id;
                rf.node_begin = (int64_t)fnodes.size();
                rf.node_count = node_count;
                for (int k = 0; k < 2; ++k) {
                    uint64_t cnum;
                    if (!hex_token(tb[node_count + k], te[node_count + k], cnum)) throw MeshError(ORC_E_IO, "invalid hex cell id");
                    rf.c[k] = cnum > 0 ? (int32_t)(cnum - 1) : -1;
                }
                for (int k = 0; k < node_count; ++k) {
                    uint64_t nn;
                    if (!hex_token(tb[k], te[k], nn)) throw MeshError(ORC_E_IO, "invalid hex node id");
                    fnodes.push_back(nn > 0 ? (int32_t)(nn - 1) : -1);
                }
                faces[idx] = rf;
                have_face[idx] = 1;
                if (!L.next(b, e)) break;
                face_number += 1;
            }
        }
        if (!L.next(hb, he)) break;
    }
    lap("sections parsed");
    m.dims = dims;
    m.n_nodes = (int64_t)have_node.size();
    for (char h : have_node) if (!h) throw MeshError(ORC_E_IO, "vertex index gap");
    m.xyz.swap(xyz);
    m.n_faces = (int64_t)have_face.size();
    for (char h : have_face) if (!h) throw MeshError(ORC_E_IO, "face index gap");
    std::sort(m.zones.begin(), m.zones.end(), [](const HostZone& a, const HostZone& b) { return a.id < b.id; });
    m.face_node_ptr.assign(m.n_faces + 1, 0);
    m.face_c0.resize(m.n_faces); m.face_c1.resize(m.n_faces); m.face_zone.resize(m.n_faces);
    for (int64_t f = 0; f < m.n_faces; ++f) m.face_node_ptr[f + 1] = m.face_node_ptr[f] + faces[f].node_count;
    m.face_nodes.resize(m.face_node_ptr[m.n_faces]);
    for (int64_t f = 0; f < m.n_faces; ++f) {
        const RawFace& rf = faces[f];
        std::copy(fnodes.begin() + rf.node_begin, fnodes.begin() + rf.node_begin + rf.node_count, m.face_nodes.begin() + m.face_node_ptr[f]);
        m.face_c0[f] = rf.c[0];
        m.face_c1[f] = rf.c[1];
        int zi = -1;
        for (size_t k = 0; k < m.zones.size(); ++k) if (m.zones[k].id == rf.zone) zi = (int)k;
        m.face_zone[f] = zi;
    }
    lap("arrays flattened");
    build_geometry(m);
    lap("geometry");
    build_derived(m);
    lap("pattern, scatter maps, level schedule");
    return mp.release();
}

HostMesh* mesh_from_arrays(int32_t dims, int64_t n_nodes, const double* xyz, int64_t n_faces, const int64_t* face_node_offsets,
                           const int64_t* face_nodes, const int64_t* c0, const int64_t* c1, const int64_t* face_zone,
                           int64_t n_zones, const int64_t* zone_ids, const int64_t* zone_types, const char* const* zone_names) {
    std::unique_ptr<HostMesh> mp(new HostMesh());
    HostMesh& m = *mp;
    m.dims = dims;
    m.n_nodes = n_nodes;
    m.n_faces = n_faces;
    m.xyz.assign(xyz, xyz + 3 * n_nodes);
    for (int64_t z = 0; z < n_zones; ++z) {
        if (!known_bc_id(zone_types[z])) throw MeshError(ORC_E_IO, "valid BC type");
        HostZone hz; hz.id = zone_ids[z]; hz.type = (int32_t)zone_types[z]; hz.name = zone_names[z];
        m.zones.push_back(hz);
    }
    std::sort(m.zones.begin(), m.zones.end(), [](const HostZone& a, const HostZone& b) { return a.id < b.id; });
    m.face_node_ptr.assign(face_node_offsets, face_node_offsets + n_faces + 1);
    m.face_nodes.resize(face_node_offsets[n_faces]);
    for (int64_t k = 0; k < face_node_offsets[n_faces]; ++k) m.face_nodes[k] = (int32_t)face_nodes[k];
    m.face_c0.resize(n_faces); m.face_c1.resize(n_faces); m.face_zone.resize(n_faces);
    // zone id -> index (ids are small integers in TGRID files; fall back to a search otherwise)
    for (int64_t f = 0; f < n_faces; ++f) {
        m.face_c0[f] = c0[f] > 0 ? (int32_t)(c0[f] - 1) : -1;
        m.face_c1[f] = c1[f] > 0 ? (int32_t)(c1[f] - 1) : -1;
        int zi = -1;
        for (size_t k = 0; k < m.zones.size(); ++k) if (m.zones[k].id == face_zone[f]) { zi = (int)k; break; }
        if (zi < 0) throw MeshError(ORC_E_IO, "face refers to an unknown zone");
        m.face_zone[f] = zi;
    }
    build_geometry(m);
    build_derived(m);
    return mp.release();
}

// The reference's Mesh flattened by the caller (orc_mesh_view): geometry as the reference computed it, no nodes.
// face_c0 / face_c1 are 0-based cell indices (cell_indices[0], cell_indices[1]; -1 = no second cell); the cell -> face lists
// must be ascending per cell like the reader leaves them (io.rs:404-411).
HostMesh* mesh_from_geometry(int32_t dims, int64_t n_cells, int64_t n_faces, const int64_t* face_c0, const int64_t* face_c1,
                             const int64_t* face_zone, const double* face_area, const double* face_normal3, const double* face_centroid3,
                             const double* cell_volume, const double* cell_centroid3, const int64_t* cell_face_offsets,
                             const int64_t* cell_face_indices, int64_t n_zones, const int64_t* zone_ids, const int64_t* zone_types,
                             const char* const* zone_names) {
    if (dims != 2 && dims != 3) throw MeshError(ORC_E_IO, "dimensions must be 2 or 3");
    if (n_cells < 0 || n_faces < 0 || n_cells >= INT32_MAX || n_faces >= INT32_MAX) throw MeshError(ORC_E_INVALID, "mesh exceeds 32-bit indexing");
    std::unique_ptr<HostMesh> mp(new HostMesh());
    HostMesh& m = *mp;
    m.dims = dims; m.n_nodes = 0; m.n_faces = n_faces; m.n_cells = n_cells;
    for (int64_t z = 0; z < n_zones; ++z) {
        if (!known_bc_id(zone_types[z])) throw MeshError(ORC_E_IO, "valid BC type");
        HostZone hz; hz.id = zone_ids[z]; hz.type = (int32_t)zone_types[z]; hz.name = zone_names[z];
        m.zones.push_back(hz);
    }
    std::sort(m.zones.begin(), m.zones.end(), [](const HostZone& a, const HostZone& b) { return a.id < b.id; });
    m.face_node_ptr.assign(n_faces + 1, 0);
    m.face_c0.resize(n_faces); m.face_c1.resize(n_faces); m.face_zone.resize(n_faces);
    for (int64_t f = 0; f < n_faces; ++f) {
        if (face_c0[f] < 0 || face_c0[f] >= n_cells || face_c1[f] < -1 || face_c1[f] >= n_cells) throw MeshError(ORC_E_INVALID, "face refers to a cell outside the mesh");
        m.face_c0[f] = (int32_t)face_c0[f];
        m.face_c1[f] = (int32_t)face_c1[f];
        int zi = -1;
        for (size_t k = 0; k < m.zones.size(); ++k) if (m.zones[k].id == face_zone[f]) { zi = (int)k; break; }
        if (zi < 0) throw MeshError(ORC_E_IO, "face refers to an unknown zone");
        m.face_zone[f] = zi;
    }
    m.face_area.assign(face_area, face_area + n_faces);
    m.face_normal.assign(face_normal3, face_normal3 + 3 * n_faces);
    m.face_centroid.assign(face_centroid3, face_centroid3 + 3 * n_faces);
    m.cell_volume.assign(cell_volume, cell_volume + n_cells);
    m.cell_centroid.assign(cell_centroid3, cell_centroid3 + 3 * n_cells);
    if (cell_face_offsets[0] != 0) throw MeshError(ORC_E_INVALID, "cell_face_offsets must start at 0");
    m.cf_ptr.resize(n_cells + 1);
    for (int64_t c = 0; c <= n_cells; ++c) {
        if (c > 0 && cell_face_offsets[c] < cell_face_offsets[c - 1]) throw MeshError(ORC_E_INVALID, "cell_face_offsets must be non-decreasing");
        if (cell_face_offsets[c] >= INT32_MAX) throw MeshError(ORC_E_INVALID, "mesh exceeds 32-bit indexing");
        m.cf_ptr[c] = (int32_t)cell_face_offsets[c];
    }
    m.cf_face.resize(m.cf_ptr[n_cells]);
    for (int64_t c = 0; c < n_cells; ++c)
        for (int32_t q = m.cf_ptr[c]; q < m.cf_ptr[c + 1]; ++q) {
            const int64_t f = cell_face_indices[q];
            if (f < 0 || f >= n_faces) throw MeshError(ORC_E_INVALID, "cell refers to a face outside the mesh");
            if (m.face_c0[f] != (int32_t)c && m.face_c1[f] != (int32_t)c) throw MeshError(ORC_E_INVALID, "cell lists a face that does not list the cell");
            if (q > m.cf_ptr[c] && f <= cell_face_indices[q - 1]) throw MeshError(ORC_E_INVALID, "cell face lists must be ascending");
            m.cf_face[q] = (int32_t)f;
        }
    build_derived(m);
    return mp.release();
}

}  // namespace orc

namespace orc {

HostMesh* extract_partition(const HostMesh& g, int32_t rank, int32_t nranks, const std::vector<int64_t>& cuts, int64_t id_offset,
                            int64_t n_global, PartPlan& plan) {
    if (nranks < 1 || rank < 0 || rank >= nranks || (int32_t)cuts.size() != nranks + 1) throw MeshError(ORC_E_INVALID, "bad rank / nranks / cuts");
    const int64_t N = g.n_cells;
    for (int32_t r = 0; r < nranks; ++r)
        if (cuts[r] > cuts[r + 1] || cuts[r] < 0 || cuts[r + 1] > N) throw MeshError(ORC_E_INVALID, "cuts must be ascending and inside the mesh");
    const int64_t g0 = cuts[rank], g1 = cuts[rank + 1];
    plan = PartPlan();
    plan.rank = rank; plan.nranks = nranks; plan.g0 = g0 + id_offset; plan.g1 = g1 + id_offset; plan.n_own = g1 - g0;
    plan.n_global = n_global;
    // halo = face neighbours of owned cells that are not owned
    std::vector<int32_t> halo;
    for (int64_t c = g0; c < g1; ++c)
        for (int32_t q = g.cf_ptr[c]; q < g.cf_ptr[c + 1]; ++q) {
            int32_t nb = g.cf_nb[q];
            if (nb >= 0 && (nb < g0 || nb >= g1)) halo.push_back(nb);
        }
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
    plan.n_lo = std::lower_bound(halo.begin(), halo.end(), (int32_t)g0) - halo.begin();
    plan.n_hi = (int64_t)halo.size() - plan.n_lo;
    const int64_t nloc = plan.n_lo + plan.n_own + plan.n_hi;
    plan.local_to_global.resize(nloc);
    std::vector<int64_t> l2w(nloc);  // local -> index in g (the window); local_to_global adds id_offset
    for (int64_t k = 0; k < plan.n_lo; ++k) l2w[k] = halo[k];
    for (int64_t k = 0; k < plan.n_own; ++k) l2w[plan.n_lo + k] = g0 + k;
    for (int64_t k = 0; k < plan.n_hi; ++k) l2w[plan.n_lo + plan.n_own + k] = halo[plan.n_lo + k];
    for (int64_t k = 0; k < nloc; ++k) plan.local_to_global[k] = l2w[k] + id_offset;
    auto to_local = [&](int64_t gc) -> int32_t {
        if (gc >= g0 && gc < g1) return (int32_t)(plan.n_lo + (gc - g0));
        auto it = std::lower_bound(halo.begin(), halo.end(), (int32_t)gc);
        int64_t k = it - halo.begin();
        return (int32_t)(k < plan.n_lo ? k : plan.n_own + k);
    };
    // exchange plan: halo cells grouped by owner (owners are contiguous global ranges, halo is sorted -> contiguous slices)
    auto owner_of = [&](int64_t gc) -> int32_t {  // last rank q with cuts[q] <= gc and a non-empty range containing gc
        int32_t q = (int32_t)(std::upper_bound(cuts.begin(), cuts.end(), gc) - cuts.begin()) - 1;
        return std::min(std::max(q, 0), nranks - 1);
    };
    plan.send_ptr.push_back(0);
    for (size_t k = 0; k < halo.size();) {
        int32_t q = owner_of(halo[k]);
        size_t e = k;
        while (e < halo.size() && owner_of(halo[e]) == q) ++e;
        plan.nbr_rank.push_back(q);
        plan.recv_begin.push_back(to_local(halo[k]));
        plan.recv_count.push_back((int32_t)(e - k));
        // what q needs from me: my owned cells adjacent to a cell owned by q (= q's halo cells that I own), ascending global id
        std::vector<int32_t> snd;
        for (size_t h = k; h < e; ++h) {
            int64_t hc = halo[h];
            for (int32_t s = g.cf_ptr[hc]; s < g.cf_ptr[hc + 1]; ++s) {
                int32_t nb = g.cf_nb[s];
                if (nb >= g0 && nb < g1) snd.push_back(nb);
            }
        }
        std::sort(snd.begin(), snd.end());
        snd.erase(std::unique(snd.begin(), snd.end()), snd.end());
        for (int32_t gc : snd) plan.send_idx.push_back(to_local(gc));
        plan.send_ptr.push_back((int32_t)plan.send_idx.size());
        k = e;
    }
    // local mesh
    std::unique_ptr<HostMesh> lp(new HostMesh());
    HostMesh& m = *lp;
    m.dims = g.dims;
    m.n_cells = nloc;
    m.zones = g.zones;
    m.zone_epoch = g.zone_epoch;
    std::vector<int32_t> faces;  // faces of owned cells, ascending global face id
    for (int64_t c = g0; c < g1; ++c)
        for (int32_t q = g.cf_ptr[c]; q < g.cf_ptr[c + 1]; ++q) faces.push_back(g.cf_face[q]);
    std::sort(faces.begin(), faces.end());
    faces.erase(std::unique(faces.begin(), faces.end()), faces.end());
    const int64_t F = (int64_t)faces.size();
    m.n_faces = F;
    m.face_c0.resize(F); m.face_c1.resize(F); m.face_zone.resize(F); m.face_area.resize(F);
    m.face_normal.resize(3 * F); m.face_centroid.resize(3 * F);
    for (int64_t k = 0; k < F; ++k) {
        const int32_t f = faces[k];
        m.face_c0[k] = to_local(g.face_c0[f]);
        m.face_c1[k] = g.face_c1[f] >= 0 ? to_local(g.face_c1[f]) : -1;
        m.face_zone[k] = g.face_zone[f];
        m.face_area[k] = g.face_area[f];
        for (int d = 0; d < 3; ++d) { m.face_normal[3 * k + d] = g.face_normal[3 * f + d]; m.face_centroid[3 * k + d] = g.face_centroid[3 * f + d]; }
    }
    m.cell_volume.resize(nloc); m.cell_centroid.resize(3 * nloc);
    for (int64_t l = 0; l < nloc; ++l) {
        const int64_t gc = l2w[l];
        m.cell_volume[l] = g.cell_volume[gc];
        for (int d = 0; d < 3; ++d) m.cell_centroid[3 * l + d] = g.cell_centroid[3 * gc + d];
    }
    // cell -> faces: owned cells keep their full list (ascending), halo cells get none
    m.cf_ptr.assign(nloc + 1, 0);
    for (int64_t c = g0; c < g1; ++c) m.cf_ptr[plan.n_lo + (c - g0) + 1] = g.cf_ptr[c + 1] - g.cf_ptr[c];
    for (int64_t l = 0; l < nloc; ++l) m.cf_ptr[l + 1] += m.cf_ptr[l];
    m.cf_face.resize(m.cf_ptr[nloc]);
    for (int64_t c = g0; c < g1; ++c) {
        int32_t o = m.cf_ptr[plan.n_lo + (c - g0)];
        for (int32_t q = g.cf_ptr[c]; q < g.cf_ptr[c + 1]; ++q)
            m.cf_face[o++] = (int32_t)(std::lower_bound(faces.begin(), faces.end(), g.cf_face[q]) - faces.begin());
    }
    m.own_lo = plan.n_lo; m.own_hi = plan.n_lo + plan.n_own;
    build_derived(m);
    return lp.release();
}

}  // namespace orc
