// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH. See orc_oracle.hpp.
// Build: g++ -std=c++17 -O2 -ffp-contract=off (Rust never contracts a*b+c into an FMA; Q19).
#include "orc_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <numeric>

namespace orc_oracle {

// =================================================================================================
// src/lib.rs:223-566 — Vector / Tensor arithmetic, one rounding per operator, left-associative.
// =================================================================================================
static inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // lib.rs:368-378
static inline Vec3 operator+(Vec3 a, Float s) { return {a.x + s, a.y + s, a.z + s}; }       // lib.rs:356-366
static inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // lib.rs:402-412
static inline Vec3 operator-(Vec3 a, Float s) { return {a.x - s, a.y - s, a.z - s}; }       // lib.rs:390-400
static inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }                         // lib.rs:529-538
static inline Vec3 operator*(Vec3 a, Float s) { return {a.x * s, a.y * s, a.z * s}; }       // lib.rs:479-508 (correct)
// lib.rs:540-549 — `impl Mul<Vector> for Float` writes rhs.y into z (quirk Q1). Reproduced on purpose.
static inline Vec3 operator*(Float s, Vec3 a) { return {a.x * s, a.y * s, a.y * s}; }
static inline Vec3 operator/(Vec3 a, Float s) { return {a.x / s, a.y / s, a.z / s}; }        // lib.rs:429-447
static inline Vec3 operator/(Vec3 a, Vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }   // lib.rs:450-459
static inline Float vdot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // lib.rs:240-242
static inline Vec3 vcross(Vec3 a, Vec3 b) {                                                  // lib.rs:254-260
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline Float vnorm(Vec3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }   // lib.rs:262-264 (powi(2) == x*x)
static inline Vec3 vunit(Vec3 a) { Float l = vnorm(a); return {a.x / l, a.y / l, a.z / l}; } // lib.rs:266-273
static inline Vec3 vabs(Vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
static inline Vec3 vones() { return {1., 1., 1.}; }
static inline Tensor3 vouter(Vec3 a, Vec3 b) {                                               // lib.rs:275-293
    return {{a.x * b.x, a.x * b.y, a.x * b.z}, {a.y * b.x, a.y * b.y, a.y * b.z}, {a.z * b.x, a.z * b.y, a.z * b.z}};
}
static inline Tensor3 operator+(Tensor3 a, Tensor3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // lib.rs:608-617
static inline Vec3 tinner(const Tensor3& t, Vec3 v) { return {vdot(t.x, v), vdot(t.y, v), vdot(t.z, v)}; }  // lib.rs:584-590

// Rust f64::min / f64::max: if one operand is NaN the other is returned (== C fmin/fmax).
static inline Float rmin(Float a, Float b) { return std::fmin(a, b); }
static inline Float rmax(Float a, Float b) { return std::fmax(a, b); }

// src/lib.rs:107-118 — TVD limiter functions psi(r)
static Float psi(int limiter, Float r) {
    switch (limiter) {
        case PSI_UD: return 0.;
        case PSI_CD1: return 1.;
        case PSI_LUD: return r;
        case PSI_QUICK: return (3. + r) / 4.;
        case PSI_UMIST: {
            Float acc = std::numeric_limits<Float>::infinity();
            acc = rmin(acc, 2. * r);
            acc = rmin(acc, (1. + 3. * r) / 4.);
            acc = rmin(acc, (3. + r) / 4.);
            acc = rmin(acc, 2.);
            return rmax(0., acc);
        }
    }
    throw Panic("unknown limiter");
}

// =================================================================================================
// nalgebra-sparse 0.9.0 / nalgebra 0.32.4 restatements (source not under /root/reference).
// =================================================================================================
size_t Csr::find(size_t i, size_t j) const {
    auto b = col.begin() + rowptr[i], e = col.begin() + rowptr[i + 1];
    auto it = std::lower_bound(b, e, j);
    if (it == e || *it != j) return NONE;
    return size_t(it - col.begin());
}
Float Csr::get(size_t i, size_t j) const {  // src/lib.rs:657-669
    size_t k = find(i, j);
    if (k == NONE) throw Panic("Tried to access CsrMatrix element that hasn't been stored yet.");
    return val[k];
}

// CsrMatrix::from(&CooMatrix): lanes sorted by minor index, duplicates combined with Add, explicit
// zeros kept (nalgebra-sparse convert::serial::convert_coo_csr). The within-lane sort is unstable in
// the crate; a stable one is used here, which only matters for the summation order of duplicates
// (on this path duplicates are 1.0 + 1.0 in restriction matrices: exact either way).
Csr coo_to_csr(const Coo& a) {
    Csr c;
    c.nrows = a.nrows; c.ncols = a.ncols;
    std::vector<size_t> counts(a.nrows + 1, 0);
    for (size_t k = 0; k < a.r.size(); ++k) counts[a.r[k] + 1]++;
    for (size_t i = 0; i < a.nrows; ++i) counts[i + 1] += counts[i];
    std::vector<size_t> pos(counts.begin(), counts.end() - 1), ucol(a.r.size());
    std::vector<Float> uval(a.r.size());
    for (size_t k = 0; k < a.r.size(); ++k) { size_t q = pos[a.r[k]]++; ucol[q] = a.c[k]; uval[q] = a.v[k]; }
    c.rowptr.assign(1, 0);
    std::vector<size_t> perm;
    for (size_t i = 0; i < a.nrows; ++i) {
        size_t b = counts[i], e = counts[i + 1];
        perm.resize(e - b);
        std::iota(perm.begin(), perm.end(), b);
        std::stable_sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return ucol[x] < ucol[y]; });
        for (size_t q = 0; q < perm.size();) {
            size_t j = ucol[perm[q]];
            Float v = uval[perm[q]];
            size_t q2 = q + 1;
            while (q2 < perm.size() && ucol[perm[q2]] == j) { v = v + uval[perm[q2]]; ++q2; }
            c.col.push_back(j); c.val.push_back(v);
            q = q2;
        }
        c.rowptr.push_back(c.col.size());
    }
    return c;
}

// &CsrMatrix * &DVector -> spmm_csr_dense(beta=0, c=zeros, alpha=1, a, b): per row
// dot = 0; for k ascending: dot += a_ik * x_k;  c_i = beta*c_i + alpha*dot.
DVec spmv(const Csr& a, const DVec& x) {
    if (x.size() != a.ncols) throw Panic("spmv: dimension mismatch");
    DVec y(a.nrows);
    for (size_t i = 0; i < a.nrows; ++i) {
        Float acc = 0.;
        for (size_t k = a.rowptr[i]; k < a.rowptr[i + 1]; ++k) acc += a.val[k] * x[a.col[k]];
        y[i] = 0. * 0. + 1. * acc;
    }
    return y;
}

// &CsrMatrix * &CsrMatrix: pattern = spmm_csr_pattern (symbolic union, sorted, nothing dropped);
// values = spmm_csr_prealloc(beta=0, c, alpha=1, a, b): c_ij = 0*c_ij; for k ascending over row i
// of a: for j over row k of b: c_ij += (alpha*a_ik) * b_kj.
Csr spgemm(const Csr& a, const Csr& b) {
    if (a.ncols != b.nrows) throw Panic("spgemm: dimension mismatch");
    Csr c;
    c.nrows = a.nrows; c.ncols = b.ncols;
    c.rowptr.assign(1, 0);
    std::vector<size_t> where(b.ncols, NONE), cols;
    for (size_t i = 0; i < a.nrows; ++i) {
        cols.clear();
        for (size_t ka = a.rowptr[i]; ka < a.rowptr[i + 1]; ++ka) {
            size_t k = a.col[ka];
            for (size_t kb = b.rowptr[k]; kb < b.rowptr[k + 1]; ++kb) {
                size_t j = b.col[kb];
                if (where[j] == NONE) { where[j] = 0; cols.push_back(j); }
            }
        }
        std::sort(cols.begin(), cols.end());
        size_t base = c.col.size();
        for (size_t q = 0; q < cols.size(); ++q) { where[cols[q]] = base + q; c.col.push_back(cols[q]); c.val.push_back(0. * 0.); }
        for (size_t ka = a.rowptr[i]; ka < a.rowptr[i + 1]; ++ka) {
            size_t k = a.col[ka];
            Float alpha_aik = 1. * a.val[ka];
            for (size_t kb = b.rowptr[k]; kb < b.rowptr[k + 1]; ++kb) c.val[where[b.col[kb]]] += alpha_aik * b.val[kb];
        }
        for (size_t j : cols) where[j] = NONE;
        c.rowptr.push_back(c.col.size());
    }
    return c;
}

// CsrMatrix::transpose(): CSR -> CSC conversion reinterpreted as the CSR of the transpose; columns sorted.
Csr transpose(const Csr& a) {
    Csr t;
    t.nrows = a.ncols; t.ncols = a.nrows;
    t.rowptr.assign(a.ncols + 1, 0);
    for (size_t k = 0; k < a.nnz(); ++k) t.rowptr[a.col[k] + 1]++;
    for (size_t j = 0; j < a.ncols; ++j) t.rowptr[j + 1] += t.rowptr[j];
    t.col.resize(a.nnz()); t.val.resize(a.nnz());
    std::vector<size_t> pos(t.rowptr.begin(), t.rowptr.end() - 1);
    for (size_t i = 0; i < a.nrows; ++i)
        for (size_t k = a.rowptr[i]; k < a.rowptr[i + 1]; ++k) { size_t q = pos[a.col[k]]++; t.col[q] = i; t.val[q] = a.val[k]; }
    return t;
}

// CsrMatrix::diagonal_as_csr(): only *stored* diagonal entries appear.
Csr diagonal_as_csr(const Csr& a) {
    Csr d;
    d.nrows = a.nrows; d.ncols = a.ncols;
    d.rowptr.assign(1, 0);
    for (size_t i = 0; i < a.nrows; ++i) {
        size_t k = a.find(i, i);
        if (k != NONE) { d.col.push_back(i); d.val.push_back(a.val[k]); }
        d.rowptr.push_back(d.col.size());
    }
    return d;
}

// nalgebra base/blas.rs `dotx`: eight interleaved accumulators over chunks of 8,
// res += acc0+acc4; res += acc1+acc5; res += acc2+acc6; res += acc3+acc7; then the tail in order.
Float dot(const DVec& a, const DVec& b) {
    if (a.size() != b.size()) throw Panic("dot: dimension mismatch");
    size_t n = a.size(), i = 0;
    Float res = 0., acc0 = 0., acc1 = 0., acc2 = 0., acc3 = 0., acc4 = 0., acc5 = 0., acc6 = 0., acc7 = 0.;
    while (n - i >= 8) {
        acc0 += a[i + 0] * b[i + 0]; acc1 += a[i + 1] * b[i + 1]; acc2 += a[i + 2] * b[i + 2]; acc3 += a[i + 3] * b[i + 3];
        acc4 += a[i + 4] * b[i + 4]; acc5 += a[i + 5] * b[i + 5]; acc6 += a[i + 6] * b[i + 6]; acc7 += a[i + 7] * b[i + 7];
        i += 8;
    }
    res += acc0 + acc4; res += acc1 + acc5; res += acc2 + acc6; res += acc3 + acc7;
    for (; i < n; ++i) res += a[i] * b[i];
    return res;
}
Float norm(const DVec& a) { return std::sqrt(dot(a, a)); }  // norm_squared().sqrt(), norm_squared = dotc(self,self)

static DVec vsub(const DVec& a, const DVec& b) { DVec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] - b[i]; return r; }
static DVec vadd(const DVec& a, const DVec& b) { DVec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i]; return r; }
static DVec vscale(Float s, const DVec& a) { DVec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = s * a[i]; return r; }

// =================================================================================================
// src/mesh.rs
// =================================================================================================
FaceZone& Mesh::get_face_zone(const std::string& name) {  // mesh.rs:189-195
    for (auto& kv : face_zones) if (kv.second.name == name) return kv.second;
    throw Panic("face zone '" + name + "' should exist in mesh");
}
Vec3 get_outward_face_normal(const Face& f, size_t cell) {  // mesh.rs:216-222
    return cell == f.cell_indices[0] ? f.normal : -f.normal;
}
static Vec3 get_inward_face_normal(const Face& f, size_t cell) { return get_outward_face_normal(f, cell) * -1.; }  // mesh.rs:224-226

static bool valid_zone_type(uint64_t t) {  // mesh.rs:50-66
    switch (t) { case 2: case 3: case 4: case 5: case 7: case 8: case 9: case 10: case 12: case 14: case 20: case 24: case 31: case 36: case 37: return true; }
    return false;
}

// =================================================================================================
// src/io.rs:289-438 — geometry pass shared by read_mesh and mesh_from_arrays.
// On entry faces hold zone, node_indices and the raw 2-entry cell_indices (NONE for "cell 0").
// =================================================================================================
static void finish_geometry(Mesh& m) {
    const int dims = m.dimensions;
    size_t max_cell = 0; bool any_cell = false;
    for (auto& f : m.faces) for (size_t c : f.cell_indices) if (c != NONE) { max_cell = std::max(max_cell, c); any_cell = true; }
    m.cells.assign(any_cell ? max_cell + 1 : 0, Cell());
    std::vector<char> seen(m.cells.size(), 0);
    for (size_t fi = 0; fi < m.faces.size(); ++fi) {
        Face& face = m.faces[fi];
        if (face.node_indices.size() < size_t(dims)) throw Panic("face has too few nodes");  // io.rs:291-294
        for (size_t n : face.node_indices) if (n >= m.vertices.size()) throw Panic("nodes should have all been read");
        auto P = [&](size_t k) -> Vec3 { return m.vertices[face.node_indices[k]]; };
        if (dims == 2) {  // io.rs:305-321
            Vec3 tangent = P(1) - P(0);
            Vec3 n = (tangent.x == 0.) ? Vec3{1., -tangent.x / tangent.y, 0.} : Vec3{-tangent.y / tangent.x, 1., 0.};
            face.normal = vunit(n);
        } else if (dims == 3) {  // io.rs:322-326
            face.normal = vunit(vcross(P(2) - P(1), P(1) - P(0)));
        } else throw Panic("dimensions must be 2 or 3");
        if (face.cell_indices.size() != 2) throw Panic("face line should carry two cell ids");
        if (face.cell_indices[0] == NONE) {  // io.rs:332-337
            face.normal = -face.normal;
            face.cell_indices.erase(face.cell_indices.begin());
        } else if (face.cell_indices[1] == NONE) {
            face.cell_indices.erase(face.cell_indices.begin() + 1);
        }
        Vec3 acc{0., 0., 0.};  // io.rs:338-342
        for (size_t k = 0; k < face.node_indices.size(); ++k) acc = acc + P(k);
        face.centroid = acc / Float(face.node_indices.size());
        size_t node_count = face.node_indices.size();
        if (node_count < 2) throw Panic("faces must have 2+ nodes");
        if (node_count == 2) {  // io.rs:345-349
            if (dims != 2) throw Panic("assertion failed: dimensions == 2");
            face.area = vnorm(P(1) - P(0));
        } else {  // io.rs:375-396: fan of triangles about the centroid
            auto tri = [](Vec3 v1, Vec3 v2, Vec3 v3) { return std::fabs(vnorm(vcross(v2 - v1, v3 - v1))) / 2.; };
            Float area = 0.;
            for (size_t k = 0; k + 1 < node_count; ++k) area = area + tri(face.centroid, P(k), P(k + 1));
            face.area = area + tri(face.centroid, P(0), P(node_count - 1));
        }
        for (size_t c : face.cell_indices) {  // io.rs:404-414
            if (c == NONE) continue;
            Cell& cell = m.cells[c];
            seen[c] = 1;
            cell.face_indices.push_back(fi);
            cell.centroid = cell.centroid + face.centroid;
        }
    }
    for (size_t ci = 0; ci < m.cells.size(); ++ci) {  // io.rs:417-438
        if (!seen[ci]) throw Panic("cell index missing from mesh (cells_hashmap.get_mut(&cell_index).unwrap())");
        Cell& cell = m.cells[ci];
        cell.centroid = cell.centroid / Float(cell.face_indices.size());
        if (cell.face_indices.size() < size_t(dims + 1)) throw Panic("cell has too few faces");
        Float vol = 0.;
        for (size_t fi : cell.face_indices) {
            const Face& f = m.faces[fi];
            vol = vol + f.area * std::fabs(vdot(f.centroid - cell.centroid, f.normal)) / Float(dims);
        }
        cell.volume = vol;
    }
}

// ---- src/io.rs:32-287 — TGRID ASCII section reader ------------------------------------------------
static std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r' || s[i] == '\f')) ++i;
        size_t b = i;
        while (i < s.size() && !(s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r' || s[i] == '\f')) ++i;
        if (i > b) out.push_back(s.substr(b, i - b));
    }
    return out;
}
static std::vector<size_t> header_items(const std::string& line) {  // io.rs:47-54: regex ([0-9a-z]+), hex
    std::vector<size_t> items;
    size_t i = 0;
    auto ok = [](char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z'); };
    while (i < line.size()) {
        while (i < line.size() && !ok(line[i])) ++i;
        size_t b = i;
        while (i < line.size() && ok(line[i])) ++i;
        if (i > b) {
            std::string tok = line.substr(b, i - b);
            size_t pos = 0;
            unsigned long long v = 0;
            try { v = std::stoull(tok, &pos, 16); } catch (...) { throw Panic("valid hex"); }
            if (pos != tok.size()) throw Panic("valid hex");
            items.push_back(size_t(v));
        }
    }
    return items;
}
static size_t parse_hex(const std::string& s) {
    size_t pos = 0;
    unsigned long long v = 0;
    try { v = std::stoull(s, &pos, 16); } catch (...) { throw Panic("invalid hex id"); }
    if (pos != s.size()) throw Panic("invalid hex id");
    return size_t(v);
}

Mesh read_mesh(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw Panic("Unable to open mesh file for reading.");
    auto next = [&](std::string& line) -> bool {
        if (!std::getline(in, line)) return false;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return true;
    };
    Mesh m;
    int dimensions = 0;
    std::string zone_name, header;
    std::map<size_t, Vec3> verts;
    std::map<size_t, Face> faces;
    if (!next(header)) throw Panic("mesh is at least one line long");
    for (;;) {
        std::vector<std::string> blocks = split_ws(header);
        const std::string tag = blocks.empty() ? std::string() : blocks[0];
        const bool zone_zero = blocks.size() > 1 && blocks[1] == "(0";  // io.rs:24-30
        if (tag == "(0") {  // io.rs:83-90
            size_t sp = header.rfind(' ');
            if (sp == std::string::npos) throw Panic("comment has a space");
            zone_name = header.substr(sp + 1);
            while (zone_name.size() >= 2 && zone_name.compare(zone_name.size() - 2, 2, "\")") == 0) zone_name.resize(zone_name.size() - 2);
        } else if (tag == "(2") {  // io.rs:92-104
            if (blocks.size() < 2 || blocks[1].empty() || blocks[1].back() != ')') throw Panic("dimensions section should have two items");
            dimensions = std::stoi(blocks[1].substr(0, blocks[1].size() - 1));
            if (dimensions != 2 && dimensions != 3) throw Panic("Mesh is not 2D or 3D.");
        } else if (tag == "(10" && !zone_zero) {  // io.rs:105-175
            std::vector<size_t> items = header_items(header);
            if (items.size() != 6) throw Panic("nodes header has six items");
            size_t node_number = items[2];
            std::string line;
            if (!next(line)) throw Panic("node section shouldn't be empty");
            for (;;) {
                if (line == "(") { if (!next(line)) throw Panic("unexpected end of node section"); continue; }
                if (!line.empty() && line[0] == ')') break;
                std::vector<std::string> lb = split_ws(line);
                if (lb.size() == size_t(dimensions)) {
                    Vec3 pos;
                    pos.x = std::stod(lb[0]); pos.y = std::stod(lb[1]); pos.z = (dimensions == 3) ? std::stod(lb[2]) : 0.;
                    verts[node_number - 1] = pos;
                }
                if (!next(line)) break;
                node_number += 1;
            }
        } else if (tag == "(12" && !zone_zero) {  // io.rs:180-193
            std::vector<size_t> items = header_items(header);
            if (items.size() != 6) throw Panic("cell section has 6 entries");
            m.cell_zones.emplace(uint64_t(items[1]), uint64_t(items[4]));
        } else if (tag == "(13" && !zone_zero) {  // io.rs:194-274
            std::vector<size_t> items = header_items(header);
            if (items.size() != 6) throw Panic("face section has 6 entries");
            size_t zone_id = items[1], start_index = items[2], boundary_type = items[4], face_type = items[5];
            if (!valid_zone_type(boundary_type)) throw Panic("valid BC type");
            if (!m.face_zones.count(zone_id)) {
                FaceZone fz; fz.zone_type = int(boundary_type); fz.name = zone_name;
                m.face_zones[zone_id] = fz;
            }
            std::string line;
            if (!next(line)) throw Panic("face section has contents");
            size_t face_number = start_index;
            for (;;) {
                if (line == "(") { if (!next(line)) throw Panic("unexpected end of face section"); continue; }
                if (!line.empty() && line[0] == ')') break;
                std::vector<std::string> lb = split_ws(line);
                if (lb.size() < 2) break;
                size_t node_count = lb.size() - 2;
                if (face_type != 0 && face_type != 5 && face_type != node_count) break;
                Face f;
                f.zone = zone_id;
                for (size_t k = node_count; k < lb.size(); ++k) { size_t c = parse_hex(lb[k]); f.cell_indices.push_back(c > 0 ? c - 1 : NONE); }
                for (size_t k = 0; k < node_count; ++k) { size_t n = parse_hex(lb[k]); f.node_indices.push_back(n > 0 ? n - 1 : NONE); }
                faces[face_number - 1] = f;
                if (!next(line)) break;
                face_number += 1;
            }
        }
        if (!next(header)) break;
    }
    m.dimensions = dimensions;
    m.vertices.resize(verts.size());
    for (size_t i = 0; i < verts.size(); ++i) {  // io.rs:498-500
        auto it = verts.find(i);
        if (it == verts.end()) throw Panic("vertex index gap");
        m.vertices[i] = it->second;
    }
    m.faces.resize(faces.size());
    for (size_t i = 0; i < faces.size(); ++i) {  // io.rs:289-290, 501-503
        auto it = faces.find(i);
        if (it == faces.end()) throw Panic("face index gap");
        m.faces[i] = it->second;
    }
    finish_geometry(m);
    return m;
}

Mesh mesh_from_arrays(int dimensions, size_t n_nodes, const Float* xyz, size_t n_faces, const int64_t* face_node_offsets,
                      const int64_t* face_nodes, const int64_t* c0, const int64_t* c1, const int64_t* face_zone,
                      size_t n_zones, const int64_t* zone_ids, const int64_t* zone_types, const char* const* zone_names) {
    Mesh m;
    m.dimensions = dimensions;
    m.vertices.resize(n_nodes);
    for (size_t i = 0; i < n_nodes; ++i) m.vertices[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    for (size_t z = 0; z < n_zones; ++z) {
        if (!valid_zone_type(uint64_t(zone_types[z]))) throw Panic("valid BC type");
        FaceZone fz; fz.zone_type = int(zone_types[z]); fz.name = zone_names[z];
        m.face_zones[uint64_t(zone_ids[z])] = fz;
    }
    m.faces.resize(n_faces);
    for (size_t f = 0; f < n_faces; ++f) {
        Face& face = m.faces[f];
        face.zone = uint64_t(face_zone[f]);
        if (!m.face_zones.count(face.zone)) throw Panic("face refers to an unknown zone");
        face.cell_indices = {c0[f] > 0 ? size_t(c0[f] - 1) : NONE, c1[f] > 0 ? size_t(c1[f] - 1) : NONE};
        for (int64_t k = face_node_offsets[f]; k < face_node_offsets[f + 1]; ++k) face.node_indices.push_back(size_t(face_nodes[k]));
    }
    finish_geometry(m);
    return m;
}

// =================================================================================================
// src/solver.rs — face / gradient helpers
// =================================================================================================
static const FaceZone& zone_of(const Mesh& m, const Face& f) {
    auto it = m.face_zones.find(f.zone);
    if (it == m.face_zones.end()) throw Panic("face zone missing");
    return it->second;
}
static inline Vec3 vel(const DVec& u, const DVec& v, const DVec& w, size_t c) { return {u[c], v[c], w[c]}; }

Vec3 get_face_velocity(const Mesh& m, const DVec& u, const DVec& v, const DVec& w, size_t fi, int interp) {  // solver.rs:952-1003
    const Face& face = m.faces[fi];
    const FaceZone& fz = zone_of(m, face);
    size_t c = face.cell_indices[0];
    switch (fz.zone_type) {
        case Wall: case VelocityInlet: return fz.vector_value;
        case PressureInlet: case PressureOutlet: case Symmetry: return vel(u, v, w, c);
        case Interior: {
            if (face.cell_indices.size() < 2) throw Panic("index out of bounds: interior face with one cell");
            size_t nb = face.cell_indices[1];
            Vec3 vel0 = vel(u, v, w, c), vel1 = vel(u, v, w, nb);
            switch (interp) {
                case V_Linear: return (vel0 + vel1) / 2.;
                case V_LinearWeighted: {
                    Float dx0 = vnorm(m.cells[c].centroid - face.centroid);
                    Float dx1 = vnorm(m.cells[nb].centroid - face.centroid);
                    return vel0 + (vel1 - vel0) * dx0 / (dx0 + dx1);
                }
                case V_RhieChow: throw Panic("unsupported");
                default: throw Panic("`None` VelocityInterpolation cannot be used for interior faces");
            }
        }
        default: throw Panic("unsupported face zone type");
    }
}

Float get_face_pressure(const Mesh& m, const DVec& p, size_t fi, int interp, int gradient) {  // solver.rs:1104-1150
    const Face& face = m.faces[fi];
    const FaceZone& fz = zone_of(m, face);
    switch (fz.zone_type) {
        case Symmetry: case Wall: case VelocityInlet: return p[face.cell_indices[0]];
        case PressureInlet: case PressureOutlet: return fz.scalar_value;
        case Interior: {
            if (face.cell_indices.size() < 2) throw Panic("index out of bounds: interior face with one cell");
            size_t c0 = face.cell_indices[0], c1 = face.cell_indices[1];
            switch (interp) {
                case P_Linear: return (p[c0] + p[c1]) * 0.5;
                case P_LinearWeighted: {
                    Float x0 = vnorm(m.cells[c0].centroid - face.centroid);
                    Float x1 = vnorm(m.cells[c1].centroid - face.centroid);
                    return p[c0] + (p[c1] - p[c0]) * x0 / (x0 + x1);
                }
                case P_Standard: throw Panic("`standard` pressure interpolation unsupported");
                case P_SecondOrder: {
                    Vec3 g0 = calculate_pressure_gradient(m, p, c0, gradient);
                    Vec3 g1 = calculate_pressure_gradient(m, p, c1, gradient);
                    Vec3 r0 = face.centroid - m.cells[c0].centroid;
                    Vec3 r1 = face.centroid - m.cells[c1].centroid;
                    return 0.5 * ((p[c0] + p[c1]) + (vdot(g0, r0) + vdot(g1, r1)));
                }
                default: throw Panic("unsupported pressure interpolation");
            }
        }
        default: throw Panic("unsupported face zone type");
    }
}

// ---- nalgebra 0.32.4 dense kernels behind the least-squares gradients (solver.rs:803-869, 903-947, 624-693) -------------------
// Restated from the crate's published source (NOT under /root/reference; "parity unpinned" like every nalgebra semantic):
//  * `&A * &x`, `&A * &B` (base/ops.rs -> blas_uninit.rs gemv_uninit / gemm_uninit): matrixmultiply is only used when every
//    dimension exceeds SMALL_DIM = 5; the 3 x n by n x 3 products here take the generic path: column j1 of the result is a
//    gemv, y = (alpha * A[:,0]) * x_0, then y = (alpha * A[:,j]) * x_j + 1 * y for j = 1.. (alpha = 1): every entry is a sum
//    accumulated in ascending inner index, one rounding per operator.
//  * `try_inverse` (linalg/inverse.rs): closed forms for 1 x 1, 2 x 2, 3 x 3 (cofactors over the determinant, `None` when the
//    determinant is exactly zero).
// Column-major storage like nalgebra's DMatrix.
struct Dense {
    size_t r = 0, c = 0;
    std::vector<Float> d;
    Dense() {}
    Dense(size_t r_, size_t c_) : r(r_), c(c_), d(r_ * c_, 0.) {}
    Float& at(size_t i, size_t j) { return d[j * r + i]; }
    Float at(size_t i, size_t j) const { return d[j * r + i]; }
};
static Dense dense_from_row_slice(size_t r, size_t c, const std::vector<Float>& data) {
    Dense a(r, c);
    for (size_t i = 0; i < r; ++i) for (size_t j = 0; j < c; ++j) a.at(i, j) = data[i * c + j];
    return a;
}
static Dense dense_transpose(const Dense& a) {
    Dense t(a.c, a.r);
    for (size_t i = 0; i < a.r; ++i) for (size_t j = 0; j < a.c; ++j) t.at(j, i) = a.at(i, j);
    return t;
}
static void dense_gemv_col(const Dense& a, const Float* x, Float* y) {  // y = A x (beta = 0, alpha = 1)
    if (a.c == 0) { for (size_t i = 0; i < a.r; ++i) y[i] = 0.; return; }
    for (size_t i = 0; i < a.r; ++i) y[i] = (1. * a.at(i, 0)) * x[0];
    for (size_t j = 1; j < a.c; ++j)
        for (size_t i = 0; i < a.r; ++i) y[i] = (1. * a.at(i, j)) * x[j] + 1. * y[i];
}
static DVec dense_mul_vec(const Dense& a, const DVec& x) { DVec y(a.r, 0.); dense_gemv_col(a, x.data(), y.data()); return y; }
static Dense dense_mul(const Dense& a, const Dense& b) {
    Dense c(a.r, b.c);
    for (size_t j1 = 0; j1 < b.c; ++j1) dense_gemv_col(a, &b.d[j1 * b.r], &c.d[j1 * c.r]);
    return c;
}
static bool dense_try_inverse(Dense& m) {  // linalg/inverse.rs try_inverse_mut, dimensions 0..3 (larger ones use LU: not reached here)
    if (m.r != m.c) throw Panic("Unable to invert a non-square matrix.");
    switch (m.r) {
        case 0: return true;
        case 1: {
            Float d = m.at(0, 0);
            if (d == 0.) return false;
            m.at(0, 0) = 1. / d;
            return true;
        }
        case 2: {
            Float m11 = m.at(0, 0), m12 = m.at(0, 1), m21 = m.at(1, 0), m22 = m.at(1, 1);
            Float det = m11 * m22 - m21 * m12;
            if (det == 0.) return false;
            m.at(0, 0) = m22 / det; m.at(0, 1) = -m12 / det;
            m.at(1, 0) = -m21 / det; m.at(1, 1) = m11 / det;
            return true;
        }
        case 3: {
            Float m11 = m.at(0, 0), m12 = m.at(0, 1), m13 = m.at(0, 2), m21 = m.at(1, 0), m22 = m.at(1, 1), m23 = m.at(1, 2),
                  m31 = m.at(2, 0), m32 = m.at(2, 1), m33 = m.at(2, 2);
            Float minor_m12_m23 = m22 * m33 - m32 * m23;
            Float minor_m11_m23 = m21 * m33 - m31 * m23;
            Float minor_m11_m22 = m21 * m32 - m31 * m22;
            Float det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
            if (det == 0.) return false;
            m.at(0, 0) = minor_m12_m23 / det;
            m.at(0, 1) = (m13 * m32 - m33 * m12) / det;
            m.at(0, 2) = (m12 * m23 - m22 * m13) / det;
            m.at(1, 0) = -minor_m11_m23 / det;
            m.at(1, 1) = (m11 * m33 - m31 * m13) / det;
            m.at(1, 2) = (m13 * m21 - m23 * m11) / det;
            m.at(2, 0) = minor_m11_m22 / det;
            m.at(2, 1) = (m12 * m31 - m32 * m11) / det;
            m.at(2, 2) = (m11 * m22 - m21 * m12) / det;
            return true;
        }
        default: throw Panic("dense_try_inverse: dimension > 3 is not on the path");
    }
}

Vec3 calculate_pressure_gradient(const Mesh& m, const DVec& p, size_t ci, int gradient) {  // solver.rs:874-949
    if (gradient == G_GreenGaussNode) throw Panic("unsupported Green-Gauss scheme");
    if (gradient == G_LeastSquares) {  // :903-947
        const Cell& cell = m.cells[ci];
        size_t n = cell.face_indices.size();
        std::vector<Float> a_data, b_data;
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            Vec3 x; Float pv;
            if (zone_of(m, face).zone_type == Interior) {
                size_t nb = face.cell_indices[0];
                if (nb == ci) nb = face.cell_indices[1];
                x = m.cells[nb].centroid - cell.centroid;
                pv = p[nb] - p[ci];
            } else {  // boundary: the face VALUE, not a difference (:927-937)
                x = face.centroid - cell.centroid;
                pv = get_face_pressure(m, p, fi, P_None, G_None);
            }
            a_data.push_back(x.x); a_data.push_back(x.y); a_data.push_back(x.z);
            b_data.push_back(pv);
        }
        Dense a = dense_from_row_slice(n, 3, a_data);
        Dense at = dense_transpose(a);
        DVec b = dense_mul_vec(at, b_data);
        Dense ata = dense_mul(at, a);
        if (!dense_try_inverse(ata)) throw Panic("called `Option::unwrap()` on a `None` value");
        DVec g = dense_mul_vec(ata, b);
        return Vec3{g[0], g[1], g[2]};
    }
    if (gradient != G_GreenGaussCell) throw Panic("unsupported gradient scheme");
    const Cell& cell = m.cells[ci];
    Vec3 acc{0., 0., 0.};
    for (size_t fi : cell.face_indices) {
        const Face& face = m.faces[fi];
        Float face_value = get_face_pressure(m, p, fi, P_Linear, gradient);
        // Float * Float * Vector: the last product is the buggy `Float * Vector` (Q1)
        Vec3 term = (face_value * (face.area / cell.volume)) * get_outward_face_normal(face, ci);
        acc = acc + term;
    }
    return acc;
}

Tensor3 calculate_velocity_gradient(const Mesh& m, const DVec& u, const DVec& v, const DVec& w, size_t ci, int gradient) {  // solver.rs:774-872
    if (gradient == G_LeastSquares) {  // :803-869
        const Cell& cell = m.cells[ci];
        size_t n = cell.face_indices.size();
        std::vector<Float> a_data, bu, bv, bw;
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            Vec3 x; Float du, dv, dw;
            if (zone_of(m, face).zone_type == Interior) {
                size_t nb = face.cell_indices[0];
                if (nb == ci) nb = face.cell_indices[1];
                x = m.cells[nb].centroid - cell.centroid;
                du = u[nb] - u[ci]; dv = v[nb] - v[ci]; dw = w[nb] - w[ci];
            } else {  // boundary: the face VALUE, not a difference (:832-838)
                Vec3 fv = get_face_velocity(m, u, v, w, fi, V_None);
                x = face.centroid - cell.centroid;
                du = fv.x; dv = fv.y; dw = fv.z;
            }
            a_data.push_back(x.x); a_data.push_back(x.y); a_data.push_back(x.z);
            bu.push_back(du); bv.push_back(dv); bw.push_back(dw);
        }
        Dense a = dense_from_row_slice(n, 3, a_data);
        Dense at = dense_transpose(a);
        DVec b_u = dense_mul_vec(at, bu), b_v = dense_mul_vec(at, bv), b_w = dense_mul_vec(at, bw);
        Dense ata = dense_mul(at, a);
        if (!dense_try_inverse(ata)) throw Panic("called `Option::unwrap()` on a `None` value");
        DVec gu = dense_mul_vec(ata, b_u), gv = dense_mul_vec(ata, b_v), gw = dense_mul_vec(ata, b_w);
        Tensor3 t;
        t.x = Vec3{gu[0], gu[1], gu[2]}; t.y = Vec3{gv[0], gv[1], gv[2]}; t.z = Vec3{gw[0], gw[1], gw[2]};
        return t;
    }
    if (gradient != G_GreenGaussCell && gradient != G_GreenGaussNode) throw Panic("unsupported gradient scheme");
    const Cell& cell = m.cells[ci];
    Tensor3 acc;
    for (size_t fi : cell.face_indices) {
        const Face& face = m.faces[fi];
        Vec3 face_value = get_face_velocity(m, u, v, w, fi, V_Linear);
        acc = acc + vouter(face_value, get_outward_face_normal(face, ci) * (face.area / cell.volume));
    }
    return acc;
}

// ---- partition emulation (NOT in the reference: the multi-GPU path's two documented deviations, SURVEY.md §8e C3/C4, restated on
// the CPU so that partitioned GPU runs can be checked against an oracle). With cuts c_0 = 0 < c_1 < ... < c_P = N set:
//  (1) momentum assembly: the diagonal of a neighbour cell that lives in ANOTHER partition is read in the state of the last
//      exchange — the value it had when the assembly started — instead of the live in-place value (j < i would see the new one);
//  (2) Multigrid: pre-smoothing and the residual are global, the coarse correction is the reference's multigrid_solve applied to
//      every partition's diagonal block of the (once scaled) fine matrix separately (aggregates never cross a cut).
struct PartitionEmu {
    std::vector<size_t> cuts;
    DVec du, dv, dw;       // diagonals at the start of the running momentum assembly
    bool lag_active = false;
    bool in_block = false;
    size_t part_of(size_t i) const { size_t r = 0; while (r + 1 < cuts.size() - 1 && i >= cuts[r + 1]) ++r; return r; }
};
static PartitionEmu g_part;
void set_partition(const std::vector<size_t>* cuts) {
    g_part = PartitionEmu();
    if (cuts && cuts->size() >= 2) g_part.cuts = *cuts;
}
static inline bool partition_on() { return g_part.cuts.size() > 2; }

// discretization.rs:14-23
static inline Float normal_momentum_coefficient(size_t i, const Csr& a_u, const Csr& a_v, const Csr& a_w, Vec3 n) {
    return vnorm(Vec3{a_u.get(i, i) * n.x, a_v.get(i, i) * n.y, a_w.get(i, i) * n.z});
}
// discretization.rs:25-34
static inline Float face_normal_momentum_coefficient(size_t i, size_t j, const Csr& a_u, const Csr& a_v, const Csr& a_w, Vec3 n) {
    return 0.5 * vnorm(Vec3{(a_u.get(i, i) + a_u.get(j, j)) * n.x, (a_v.get(i, i) + a_v.get(j, j)) * n.y, (a_w.get(i, i) + a_w.get(j, j)) * n.z});
}

Float get_face_flux(const Mesh& m, const DVec& u, const DVec& v, const DVec& w, const DVec& p, size_t fi, size_t ci,
                    int vel_interp, int gradient, const Csr& a_u, const Csr& a_v, const Csr& a_w) {  // solver.rs:1007-1102
    const Face& face = m.faces[fi];
    Vec3 n = get_outward_face_normal(face, ci);
    const FaceZone& fz = zone_of(m, face);
    switch (fz.zone_type) {
        case Wall: case Symmetry: return 0.;
        case VelocityInlet: case PressureInlet: case PressureOutlet:
            return vdot(n, get_face_velocity(m, u, v, w, fi, V_None));
        case Interior:
            switch (vel_interp) {
                case V_Linear: case V_LinearWeighted: return vdot(n, get_face_velocity(m, u, v, w, fi, vel_interp));
                case V_RhieChow: {
                    size_t nb = face.cell_indices[0];
                    if (nb == ci) {
                        if (face.cell_indices.size() < 2) throw Panic("index out of bounds: interior face with one cell");
                        nb = face.cell_indices[1];
                    }
                    Vec3 vel_i = vel(u, v, w, ci), vel_j = vel(u, v, w, nb);
                    Vec3 d = m.cells[nb].centroid - m.cells[ci].centroid;
                    Float a_i = normal_momentum_coefficient(ci, a_u, a_v, a_w, n);
                    Float a_j = (g_part.lag_active && g_part.part_of(nb) != g_part.part_of(ci))
                                    ? vnorm(Vec3{g_part.du[nb] * n.x, g_part.dv[nb] * n.y, g_part.dw[nb] * n.z})   // partition emulation (1)
                                    : normal_momentum_coefficient(nb, a_u, a_v, a_w, n);
                    Vec3 g_i = calculate_pressure_gradient(m, p, ci, gradient);
                    Vec3 g_j = calculate_pressure_gradient(m, p, nb, gradient);
                    Float vol_i = m.cells[ci].volume, vol_j = m.cells[nb].volume;
                    Float term_1 = vdot(vel_i + vel_j, n);
                    Float term_2 = (vol_i / a_i + vol_j / a_j) * (p[ci] - p[nb]) / vnorm(d);
                    Float term_3 = vdot((vol_i / a_i) * g_i + (vol_j / a_j) * g_j, vunit(d));  // Float * Vector: Q1
                    return 0.5 * (term_1 + term_2 - term_3);
                }
                default: throw Panic("`None` VelocityInterpolation cannot be used for interior faces");
            }
        default: throw Panic("unsupported face zone type");
    }
}

// =================================================================================================
// src/discretization.rs
// =================================================================================================
void build_momentum_diffusion_matrix(const Mesh& m, Float mu, Csr& a_out, DVec& b_u, DVec& b_v, DVec& b_w) {  // :39-131
    size_t n = m.cells.size();
    Coo a; a.nrows = a.ncols = n;
    b_u.assign(n, 0.); b_v.assign(n, 0.); b_w.assign(n, 0.);
    for (size_t ci = 0; ci < n; ++ci) {
        const Cell& cell = m.cells[ci];
        Float a_p = 0.;
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            const FaceZone& bc = zone_of(m, face);
            Float d_i; size_t nb;
            switch (bc.zone_type) {
                case Wall: case VelocityInlet: {
                    d_i = mu * face.area / vnorm(face.centroid - cell.centroid);
                    Vec3 src = bc.vector_value * d_i;
                    b_u[ci] += src.x; b_v[ci] += src.y; b_w[ci] += src.z;
                    nb = NONE;
                    break;
                }
                case PressureInlet: case PressureOutlet: case Symmetry: d_i = 0.; nb = NONE; break;
                case Interior: {
                    nb = face.cell_indices[0];
                    if (nb == ci) {
                        if (face.cell_indices.size() < 2) throw Panic("interior faces should have two neighbors");
                        nb = face.cell_indices[1];
                    }
                    Vec3 e_xi = m.cells[nb].centroid - cell.centroid;
                    d_i = mu * face.area / vnorm(e_xi);
                    break;
                }
                default: throw Panic("BC not supported");
            }
            a_p += d_i;
            if (nb != NONE) a.push(ci, nb, -d_i);
        }
        a.push(ci, ci, a_p);
    }
    a_out = coo_to_csr(a);
}

Csr initialize_momentum_matrix(const Mesh& m) {  // :450-472
    size_t n = m.cells.size();
    Coo a; a.nrows = a.ncols = n;
    for (size_t ci = 0; ci < n; ++ci) {
        const Cell& cell = m.cells[ci];
        a.push(ci, ci, 1.);
        Float nf = Float(cell.face_indices.size());
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            if (face.cell_indices.size() == 2) {
                size_t nb = face.cell_indices[0] == ci ? face.cell_indices[1] : face.cell_indices[0];
                a.push(ci, nb, -1. / nf);
            }
        }
    }
    return coo_to_csr(a);
}

static inline Float total_max(Float a, Float b) {  // max_by(total_cmp) over [a, b]: last maximum wins; NaN ordering by total_cmp
    auto key = [](Float x) { int64_t i; std::memcpy(&i, &x, 8); return i ^ int64_t(uint64_t(i >> 63) >> 1); };
    return key(b) >= key(a) ? b : a;
}
static inline Float total_min(Float a, Float b) {  // min_by(total_cmp): first minimum wins
    auto key = [](Float x) { int64_t i; std::memcpy(&i, &x, 8); return i ^ int64_t(uint64_t(i >> 63) >> 1); };
    return key(b) < key(a) ? b : a;
}

Peclet build_momentum_advection_matrices(Csr& a_u, Csr& a_v, Csr& a_w, DVec& b_u, DVec& b_v, DVec& b_w, const Csr& a_di,
                                         const Mesh& m, const DVec& u, const DVec& v, const DVec& w, const DVec& p,
                                         int momentum, int limiter, int vel_interp, int p_interp, int gradient, Float rho) {  // :134-356
    Float min_pe = std::numeric_limits<Float>::infinity(), max_pe = -std::numeric_limits<Float>::infinity(), avg_pe = 0.;
    size_t n = m.cells.size();
    struct LagScope {  // partition emulation (1): snapshot of the diagonals = what the last halo exchange delivered
        bool on;
        LagScope(const Csr& a_u, const Csr& a_v, const Csr& a_w, size_t n) : on(partition_on() && n == g_part.cuts.back()) {
            if (!on) return;
            g_part.du.resize(n); g_part.dv.resize(n); g_part.dw.resize(n);
            for (size_t i = 0; i < n; ++i) { g_part.du[i] = a_u.get(i, i); g_part.dv[i] = a_v.get(i, i); g_part.dw[i] = a_w.get(i, i); }
            g_part.lag_active = true;
        }
        ~LagScope() { g_part.lag_active = false; }
    } lag_scope(a_u, a_v, a_w, n);
    for (size_t ci = 0; ci < n; ++ci) {
        const Cell& cell = m.cells[ci];
        Vec3 s_u{0., 0., 0.};  // get_momentum_source_term == 0 (solver.rs:698-701)
        Vec3 s_u_dc{0., 0., 0.}, s_d_cross{0., 0., 0.};
        Float a_ii_di = a_di.get(ci, ci);
        Vec3 a_p{0., 0., 0.};
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            Float face_flux = get_face_flux(m, u, v, w, p, fi, ci, vel_interp, gradient, a_u, a_v, a_w);
            Vec3 n_out = get_outward_face_normal(face, ci);
            Float f_i = face_flux * face.area * rho;
            Float face_pressure = get_face_pressure(m, p, fi, p_interp, gradient);
            size_t nb = face.cell_indices.size() == 1 ? NONE : (face.cell_indices[0] == ci ? face.cell_indices[1] : face.cell_indices[0]);
            Vec3 a_nb;
            switch (momentum) {
                case UD: a_nb = rmin(f_i, 0.) * vones(); break;
                case CD1: a_nb = f_i * vones() / 2.; break;
                case TVD: {
                    if (nb == NONE) {
                        a_nb = rmin(f_i, 0.) * vones();
                    } else {
                        size_t downstream = f_i > 0. ? nb : ci;
                        Vec3 dvel = vel(u, v, w, downstream), cvel = vel(u, v, w, ci);
                        if (vnorm(dvel - cvel) == 0.) {
                            a_nb = f_i * vones() / 2.;
                        } else {
                            Tensor3 grad = calculate_velocity_gradient(m, u, v, w, ci, gradient);
                            Vec3 r_pa = m.cells[nb].centroid - m.cells[ci].centroid;
                            Vec3 r = 2. * tinner(grad, r_pa) / (dvel - cvel) - 1.;           // `2. * Vector`: Q1
                            a_nb = f_i * Vec3{psi(limiter, r.x), psi(limiter, r.y), psi(limiter, r.z)} / 2.;  // Q1 again
                        }
                    }
                    break;
                }
                default: throw Panic("unsupported momentum scheme");
            }
            a_p = a_p + (-a_nb + f_i);
            s_u = s_u + (-n_out) * face_pressure * face.area;
            if (nb == NONE) {
                const FaceZone& fz = zone_of(m, face);
                if (fz.zone_type == Wall || fz.zone_type == VelocityInlet)
                    s_u = s_u + Vec3{(a_nb.x - f_i) * fz.vector_value.x, (a_nb.y - f_i) * fz.vector_value.y, (a_nb.z - f_i) * fz.vector_value.z};
                else
                    s_u = s_u + Vec3{0., 0., 0.};
            } else {
                Float a_ij_di = a_di.get(ci, nb);
                size_t ku = a_u.find(ci, nb), kv = a_v.find(ci, nb), kw = a_w.find(ci, nb);
                if (ku == NONE || kv == NONE || kw == NONE) throw Panic("momentum matrix entry not stored");
                a_u.val[ku] = a_nb.x + a_ij_di;
                a_v.val[kv] = a_nb.y + a_ij_di;
                a_w.val[kw] = a_nb.z + a_ij_di;
            }
        }
        Vec3 source_total = s_u + s_u_dc + s_d_cross;
        b_u[ci] = source_total.x; b_v[ci] = source_total.y; b_w[ci] = source_total.z;
        Float pe_x = a_p.x / a_ii_di, pe_y = a_p.y / a_ii_di, pe_z = a_p.z / a_ii_di;
        max_pe = total_max(total_max(total_max(max_pe, pe_x), pe_y), pe_z);
        min_pe = total_min(total_min(total_min(min_pe, pe_x), pe_y), pe_z);
        avg_pe += (((0. + pe_x) + pe_y) + pe_z) / 3.;
        size_t du = a_u.find(ci, ci), dv = a_v.find(ci, ci), dw = a_w.find(ci, ci);
        if (du == NONE || dv == NONE || dw == NONE) throw Panic("momentum matrix diagonal not stored");
        a_u.val[du] = a_p.x + a_ii_di;
        a_v.val[dv] = a_p.y + a_ii_di;
        a_w.val[dw] = a_p.z + a_ii_di;
    }
    return {avg_pe / Float(n), min_pe, max_pe};
}

void build_pressure_correction_matrices(const Mesh& m, const DVec& u, const DVec& v, const DVec& w, const DVec& p,
                                        const Csr& a_u, const Csr& a_v, const Csr& a_w, const NumericalSettings& s, Float rho,
                                        Csr& a_out, DVec& b) {  // :359-448
    size_t n = m.cells.size();
    Coo a; a.nrows = a.ncols = n;
    b.assign(n, 0.);
    for (size_t ci = 0; ci < n; ++ci) {
        const Cell& cell = m.cells[ci];
        Float a_p = 0., b_p = 0.;
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            Float flux = get_face_flux(m, u, v, w, p, fi, ci, s.velocity_interpolation, s.gradient_reconstruction, a_u, a_v, a_w);
            Vec3 n_in = get_inward_face_normal(face, ci);
            b_p += rho * (-flux) * face.area;
            if (face.cell_indices.size() > 1) {
                size_t nb = face.cell_indices[0] != ci ? face.cell_indices[0] : face.cell_indices[1];
                Float a_mag = face_normal_momentum_coefficient(ci, nb, a_u, a_v, a_w, n_in);
                Float a_nb = rho * (face.area * face.area) / a_mag;
                a.push(ci, nb, -a_nb);
                a_p += a_nb;
            } else {
                Float a_ii_norm = normal_momentum_coefficient(ci, a_u, a_v, a_w, n_in);
                Float a_nb = rho * (face.area * face.area) / a_ii_norm;
                a_p += a_nb / 2.;
            }
        }
        a.push(ci, ci, a_p);
        b[ci] = b_p;
    }
    a_out = coo_to_csr(a);
}

// =================================================================================================
// src/linear_algebra.rs
// =================================================================================================
static MgTrace* g_trace = nullptr;
void set_mg_trace(MgTrace* t) { g_trace = t; }

Csr build_restriction_matrix(const Csr& a, int method) {  // :12-63
    size_t n = a.ncols / 2 + a.ncols % 2;
    Coo r; r.nrows = n; r.ncols = a.ncols;
    if (method == Injection) {
        for (size_t row = 0; row + 1 < n; ++row) { r.push(row, 2 * row, 1.); r.push(row, 2 * row + 1, 1.); }
        r.push(n - 1, 2 * (n - 1), 1.);
        if (2 * (n - 1) + 1 < a.ncols) r.push(n - 1, 2 * (n - 1) + 1, 1.);
    } else {
        std::vector<char> combined(a.ncols, 0);  // the HashSet is only probed/inserted: order-independent
        for (size_t i = 0; i < a.nrows; ++i) {
            Float strongest = std::numeric_limits<Float>::max();
            size_t pick = NONE;
            for (size_t k = a.rowptr[i]; k < a.rowptr[i + 1]; ++k) {
                size_t j = a.col[k];
                if (combined[j] || i == j) continue;
                Float coeff = a.val[k];
                if (coeff < strongest) { strongest = coeff; pick = j; }
            }
            if (pick != NONE) {
                combined[pick] = 1;
                r.push(i / 2, i, 1.0);
                r.push(i / 2, pick, 1.0);
            }
        }
    }
    return coo_to_csr(r);
}

static void bicgstab(const Csr& a, const DVec& b, DVec& x, uint64_t iterations) {  // :247-269
    size_t n = b.size();
    DVec r = vsub(b, spmv(a, x));
    DVec r_hat_0(n, 1.);
    Float rho = dot(r, r_hat_0);
    DVec p = r;
    for (uint64_t it = 0; it < iterations; ++it) {
        DVec nu = spmv(a, p);
        Float alpha = rho / dot(r_hat_0, nu);
        DVec h = vadd(x, vscale(alpha, p));
        DVec s = vsub(r, vscale(alpha, nu));
        DVec t = spmv(a, s);
        Float omega = dot(t, s) / dot(t, t);
        x = vadd(h, vscale(omega, s));
        r = vsub(s, vscale(omega, t));
        Float rho_prev = rho;
        rho = dot(r_hat_0, r);
        Float beta = rho / rho_prev * alpha / omega;
        p = vadd(r, vscale(beta, vsub(p, vscale(omega, nu))));
    }
}

static DVec multigrid_solve(const Csr& a, const DVec& r, uint64_t level, uint64_t max_levels, int smooth_method, uint64_t smooth_iters,
                            Float smooth_relax, Float smooth_thr, int restriction, int preconditioner, const SolveOpts& o) {  // :66-141
    Csr R = build_restriction_matrix(a, restriction);
    DVec r_prime = spmv(R, r);
    Csr Rt = transpose(R);
    Csr a_prime = spgemm(spgemm(R, a), Rt);
    if (g_trace) { g_trace->restriction.push_back(R); g_trace->coarse.push_back(a_prime); }
    DVec e_prime(a_prime.ncols, 0.);
    iterative_solve(a_prime, r_prime, e_prime, smooth_iters, smooth_method, smooth_relax, smooth_thr, preconditioner, o);
    Float error_magnitude = norm(vsub(r_prime, spmv(a_prime, e_prime)));
    if (std::isnan(error_magnitude)) throw Panic("Multigrid diverged");
    if (level < max_levels && a_prime.nrows > 16) {
        DVec corr = multigrid_solve(a_prime, r_prime, level + 1, max_levels, smooth_method, smooth_iters, smooth_relax, smooth_thr,
                                    restriction, preconditioner, o);
        for (size_t i = 0; i < e_prime.size(); ++i) e_prime[i] += corr[i];
        iterative_solve(a_prime, r_prime, e_prime, smooth_iters, smooth_method, smooth_relax, smooth_thr / 10., preconditioner, o);
    }
    return spmv(transpose(R), e_prime);
}

void iterative_solve(const Csr& a, const DVec& b, DVec& x, uint64_t iterations, int method, Float relaxation, Float threshold,
                     int preconditioner, const SolveOpts& o) {  // :144-299
    Csr a_tmp; DVec b_tmp;
    const Csr* ap = &a; const DVec* bp = &b;
    if (preconditioner == PC_Jacobi) {  // :159-167
        Csr p_inv = diagonal_as_csr(a);
        for (Float& v : p_inv.val) v = 1. / v;
        a_tmp = spgemm(p_inv, a);
        b_tmp = spmv(p_inv, b);
        ap = &a_tmp; bp = &b_tmp;
    }
    const Csr& A = *ap; const DVec& B = *bp;
    Float initial_residual = 0.;
    switch (method) {
        case Jacobi: {  // :172-218
            Csr a_prime = A;
            for (size_t i = 0; i < A.nrows; ++i)
                for (size_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
                    a_prime.val[k] = (A.col[k] == i) ? 0. : A.val[k] / A.get(i, i);
            DVec b_prime(B.size());
            for (size_t i = 0; i < B.size(); ++i) b_prime[i] = B[i] / A.get(i, i);
            for (uint64_t it = 0; it < iterations; ++it) {
                for (Float v : x) if (std::isnan(v)) throw Panic("diverged");
                DVec ax = spmv(a_prime, x);
                DVec nx(x.size());
                Float om = 1. - relaxation;
                for (size_t i = 0; i < x.size(); ++i) nx[i] = relaxation * (b_prime[i] - ax[i]) + x[i] * om;
                x = nx;
                Float r = norm(vsub(B, spmv(A, x)));
                if (x.empty()) throw Panic("called `Option::unwrap()` on a `None` value");
                Float max_abs = std::fabs(x[0]);  // max_by(|a|.total_cmp(|b|)).abs()
                for (size_t i = 1; i < x.size(); ++i) max_abs = total_max(max_abs, std::fabs(x[i]));
                if (it == 1) initial_residual = r;
                else if (r / initial_residual < threshold) break;
                if (max_abs > 1e10) throw Panic("Diverged - max solution value > 10^10");
            }
            break;
        }
        case GaussSeidel: {  // :219-246
            if (!o.gs_intended) {
                // As written: `a.get(i, j)` over every j panics on the first un-stored entry, and the arm
                // ends in an unconditional panic!().
                throw Panic("Gauss-Seidel out for maintenance :)");
            }
            for (uint64_t it = 0; it < iterations; ++it) {
                for (size_t i = 0; i < A.nrows; ++i) {
                    Float sum = 0.;
                    for (size_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
                        if (A.col[k] != i) sum += A.val[k] * x[A.col[k]];
                    x[i] = x[i] * (1. - relaxation) + relaxation * (B[i] - sum) / A.get(i, i);
                    if (std::isnan(x[i])) throw Panic("****** Solution diverged ******");
                }
            }
            break;
        }
        case BiCGSTAB: bicgstab(A, B, x, iterations); break;
        case Multigrid: {  // :270-296
            iterative_solve(A, B, x, iterations, o.mg_smoother, relaxation, threshold, preconditioner, o);
            DVec r = vsub(B, spmv(A, x));
            if (partition_on() && !g_part.in_block && A.nrows == g_part.cuts.back()) {  // partition emulation (2)
                g_part.in_block = true;
                try {
                    for (size_t q = 0; q + 1 < g_part.cuts.size(); ++q) {
                        const size_t lo = g_part.cuts[q], hi = g_part.cuts[q + 1];
                        Csr blk; blk.nrows = blk.ncols = hi - lo; blk.rowptr.assign(1, 0);
                        for (size_t i = lo; i < hi; ++i) {
                            for (size_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
                                if (A.col[k] >= lo && A.col[k] < hi) { blk.col.push_back(A.col[k] - lo); blk.val.push_back(A.val[k]); }
                            blk.rowptr.push_back(blk.col.size());
                        }
                        DVec rb(r.begin() + lo, r.begin() + hi);
                        DVec corr = multigrid_solve(blk, rb, 1, o.mg_levels, o.mg_smoother, iterations, relaxation, threshold, Strongest, preconditioner, o);
                        for (size_t i = lo; i < hi; ++i) x[i] += corr[i - lo];
                    }
                } catch (...) { g_part.in_block = false; throw; }
                g_part.in_block = false;
                break;
            }
            DVec corr = multigrid_solve(A, r, 1, o.mg_levels, o.mg_smoother, iterations, relaxation, threshold, Strongest, preconditioner, o);
            for (size_t i = 0; i < x.size(); ++i) x[i] += corr[i];
            break;
        }
        default: throw Panic("unsupported solution method");
    }
}

// =================================================================================================
// src/solver.rs — correction + SIMPLE driver
// =================================================================================================
CorrectionNorms apply_pressure_correction(const Mesh& m, const Csr& a_u, const Csr& a_v, const Csr& a_w, const DVec& p_prime,
                                          DVec& u, DVec& v, DVec& w, DVec& p, const NumericalSettings& s) {  // :1170-1227
    Float velocity_corr_sum = 0.;
    for (size_t ci = 0; ci < m.cells.size(); ++ci) {
        const Cell& cell = m.cells[ci];
        Float pc = ci < p_prime.size() ? p_prime[ci] : 0.;
        p[ci] += s.pressure_relaxation * pc;
        Vec3 acc{0., 0., 0.};
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            const FaceZone& fz = zone_of(m, face);
            Vec3 n = get_outward_face_normal(face, ci);
            Float pn;
            switch (fz.zone_type) {
                case Wall: case Symmetry: case VelocityInlet: pn = p_prime[ci]; break;
                case PressureInlet: case PressureOutlet: pn = 0.; break;
                case Interior: pn = p_prime[face.cell_indices[0] == ci ? face.cell_indices[1] : face.cell_indices[0]]; break;
                default: throw Panic("BC not supported");
            }
            Vec3 scaled{n.x / a_u.get(ci, ci), n.y / a_v.get(ci, ci), n.z / a_w.get(ci, ci)};
            acc = acc + scaled * (p_prime[ci] - pn) * face.area;
        }
        u[ci] += acc.x * s.momentum_relaxation;
        v[ci] += acc.y * s.momentum_relaxation;
        w[ci] += acc.z * s.momentum_relaxation;
        Float nn = vnorm(acc);
        velocity_corr_sum += nn * nn;
    }
    return {norm(p_prime), std::sqrt(velocity_corr_sum)};
}

static Float seq_sum(const DVec& a) { Float s = 0.; for (Float v : a) s += v; return s; }

void solve_steady(Mesh& m, DVec& u, DVec& v, DVec& w, DVec& p, const NumericalSettings& s, Float rho, Float mu,
                  uint64_t iteration_count, uint64_t reporting_interval, ReportFn cb, void* user, PhaseTimes* times) {  // :26-244
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    size_t n = m.cells.size();
    Csr a_di; DVec b_u_di, b_v_di, b_w_di;
    build_momentum_diffusion_matrix(m, mu, a_di, b_u_di, b_v_di, b_w_di);
    Csr a_u = initialize_momentum_matrix(m), a_v = initialize_momentum_matrix(m), a_w = initialize_momentum_matrix(m);
    DVec b_u(n, 0.), b_v(n, 0.), b_w(n, 0.), p_prime(n, 0.);
    const MatrixSolverSettings& ms = s.matrix_solver;
    for (uint64_t iter = 1; iter <= iteration_count; ++iter) {
        auto t0 = clk::now();
        Peclet pe = build_momentum_advection_matrices(a_u, a_v, a_w, b_u, b_v, b_w, a_di, m, u, v, w, p, s.momentum, s.limiter,
                                                      s.velocity_interpolation, s.pressure_interpolation, s.gradient_reconstruction, rho);
        for (size_t i = 0; i < n; ++i) { b_u[i] += b_u_di[i]; b_v[i] += b_v_di[i]; b_w[i] += b_w_di[i]; }
        auto t1 = clk::now();
        iterative_solve(a_u, b_u, u, ms.iterations, ms.solver_type, ms.relaxation, ms.relative_convergence_threshold, ms.preconditioner, s.opts);
        iterative_solve(a_v, b_v, v, ms.iterations, ms.solver_type, ms.relaxation, ms.relative_convergence_threshold, ms.preconditioner, s.opts);
        iterative_solve(a_w, b_w, w, ms.iterations, ms.solver_type, ms.relaxation, ms.relative_convergence_threshold, ms.preconditioner, s.opts);
        auto t2 = clk::now();
        Csr pc_a; DVec pc_b;
        build_pressure_correction_matrices(m, u, v, w, p, a_u, a_v, a_w, s, rho, pc_a, pc_b);
        auto t3 = clk::now();
        for (Float& x : p_prime) x *= 0.;
        iterative_solve(pc_a, pc_b, p_prime, ms.iterations, ms.solver_type, ms.relaxation, ms.relative_convergence_threshold, ms.preconditioner, s.opts);
        auto t4 = clk::now();
        CorrectionNorms cn = apply_pressure_correction(m, a_u, a_v, a_w, p_prime, u, v, w, p, s);
        Float u_avg = seq_sum(u) / Float(n), v_avg = seq_sum(v) / Float(n), w_avg = seq_sum(w) / Float(n);
        auto t5 = clk::now();
        if (times) {
            times->momentum_assembly += secs(t0, t1); times->momentum_solves += secs(t1, t2); times->pressure_assembly += secs(t2, t3);
            times->pressure_solve += secs(t3, t4); times->correction += secs(t4, t5);
        }
        if (cb && reporting_interval != 0 && iter % reporting_interval == 0) {
            IterationReport rep{iter, u_avg, v_avg, w_avg, pe.avg, pe.min, pe.max, cn.velocity_corr, cn.p_prime_norm};
            cb(&rep, user);
        }
        if (std::isnan(u_avg) || std::isnan(v_avg) || std::isnan(w_avg)) throw Panic("solution diverged");
    }
    // :227-242 — mean |grad p|, |grad u| are computed and discarded by the reference (Q15); nothing to restate.
}


// =================================================================================================
// src/solver.rs:246-352, 414-509, 703-770 — flow initialisation (the step before the SIMPLE loop)
// =================================================================================================
static inline Vec3 vreciprocal(Vec3 a) {  // lib.rs:244-252: zero components stay zero
    return {a.x != 0. ? 1. / a.x : 0., a.y != 0. ? 1. / a.y : 0., a.z != 0. ? 1. / a.z : 0.};
}
static inline Float vector_angle(Vec3 a, Vec3 b) { return std::acos(vdot(a, b) / (vnorm(a) * vnorm(b))); }  // lib.rs:645-647

int check_boundary_conditions(const Mesh& m) {  // :710-770
    const Float PI = Float(3.14159274101257324f);  // std::f32::consts::PI as Float
    const Float TOL = 5. * 180. / PI;              // "5 degrees" (:713): 286.5 rad — neither check below can ever fire
    uint16_t pressure_bc_count = 0, velocity_bc_count = 0;  // u16 counters: a release build wraps (a debug build would panic
                                                            // on overflow for meshes with > 65535 faces)
    for (const auto& kv : m.face_zones) {  // HashMap order in the reference; the result does not depend on it
        const FaceZone& z = kv.second;
        switch (z.zone_type) {
            case Wall:
                if (vnorm(z.vector_value) > 0.) {
                    for (const Face& face : m.faces) {  // every face of the MESH, not of the zone (:722)
                        velocity_bc_count = uint16_t(velocity_bc_count + 1);
                        if (PI / 2. - std::fabs(vector_angle(face.normal, z.vector_value)) > TOL) throw Panic("Wall velocity must be tangent to faces in zone.");
                    }
                }
                break;
            case VelocityInlet:
                velocity_bc_count = uint16_t(velocity_bc_count + 1);
                for (const Face& face : m.faces)
                    if (std::fabs(vector_angle(face.normal, z.vector_value)) > TOL) throw Panic("VelocityInlet velocity must not be tangent to faces in zone.");
                break;
            case PressureInlet: case PressureOutlet: pressure_bc_count = uint16_t(pressure_bc_count + 1); break;
            default: break;
        }
    }
    if (velocity_bc_count > 0) {
        if (pressure_bc_count > 1) return Hybrid;
        return VelocityOnly;
    }
    if (pressure_bc_count > 0) return PressureOnly;
    throw Panic("You must set boundary conditions.");
}

void build_pressure_laplace(const Mesh& m, Csr& a_out, DVec& b) {  // :437-494
    size_t n = m.cells.size();
    Coo a; a.nrows = a.ncols = n;
    b.assign(n, 0.);
    for (size_t ci = 0; ci < n; ++ci) {
        const Cell& cell = m.cells[ci];
        Float a_p = 0.;
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            Vec3 nout = get_outward_face_normal(face, ci);
            const FaceZone& z = zone_of(m, face);
            Float a_nb, source; size_t nb;
            switch (z.zone_type) {
                case Interior:
                    nb = face.cell_indices[0] == ci ? face.cell_indices[1] : face.cell_indices[0];
                    a_nb = vdot(vreciprocal(cell.centroid - m.cells[nb].centroid), nout) * (face.area / cell.volume);
                    source = 0.;
                    break;
                case PressureInlet: case PressureOutlet:
                    a_nb = vdot(vreciprocal(cell.centroid - face.centroid), nout) * (face.area / cell.volume);
                    source = a_nb * z.scalar_value;
                    nb = NONE;
                    break;
                default: a_nb = 0.; source = 0.; nb = NONE;  // Symmetry | Wall | everything else (:476-485)
            }
            if (nb != NONE) a.push(ci, nb, -a_nb);
            b[ci] += source;
            a_p += a_nb;
        }
        a.push(ci, ci, a_p);
    }
    a_out = coo_to_csr(a);
}

void initialize_pressure_field(const Mesh& m, DVec& p) {  // :414-509
    Csr a; DVec b;
    build_pressure_laplace(m, a, b);
    iterative_solve(a, b, p, 10, Jacobi, 0.1, 1e-6, PC_Jacobi);  // :498-507
}

// `&a * sa + &b * sb` (:310-311). nalgebra-sparse 0.9.0: `&Csr * scalar` maps every value to v * scalar; `&a + &b` builds the
// union pattern with zero values and runs spadd_csr_prealloc twice: c = 0 * c + 1 * a, then c = 1 * c + 1 * b.
Csr csr_blend(const Csr& a, Float sa, const Csr& b, Float sb) {
    if (a.rowptr != b.rowptr || a.col != b.col) throw Panic("csr_blend: the restatement covers matrices with one pattern (all mesh matrices share it)");
    Csr c = a;
    for (size_t k = 0; k < c.val.size(); ++k) {
        Float v = 0.;
        v = v + 1. * (a.val[k] * sa);
        v = v + 1. * (b.val[k] * sb);
        c.val[k] = v;
    }
    return c;
}

void initialize_flow(const Mesh& m, Float mu, Float rho, uint64_t iteration_count, DVec& u, DVec& v, DVec& w, DVec& p) {  // :246-352
    check_boundary_conditions(m);
    size_t n = m.cells.size();
    u.assign(n, 0.); v.assign(n, 0.); w.assign(n, 0.); p.assign(n, 0.);
    Csr a_di; DVec b_u_di, b_v_di, b_w_di;
    build_momentum_diffusion_matrix(m, mu, a_di, b_u_di, b_v_di, b_w_di);
    Csr a_u = initialize_momentum_matrix(m), a_v = initialize_momentum_matrix(m), a_w = initialize_momentum_matrix(m);
    DVec b_u(n, 0.), b_v(n, 0.), b_w(n, 0.);
    initialize_pressure_field(m, p);
    build_momentum_advection_matrices(a_u, a_v, a_w, b_u, b_v, b_w, a_di, m, u, v, w, p, UD, PSI_UD, V_LinearWeighted, P_LinearWeighted,
                                      G_GreenGaussCell, rho);  // :288-306
    for (size_t i = 0; i < n; ++i) { b_u[i] += b_u_di[i]; b_v[i] += b_v_di[i]; b_w[i] += b_w_di[i]; }
    Float diffusion_fraction = 1.;
    while (diffusion_fraction >= 0.) {  // 1, 0.8, 0.6000000000000001, 0.4000000000000001, 0.20000000000000007, 5.6e-17: six rounds
        iterative_solve(csr_blend(a_u, 1. - diffusion_fraction, a_di, diffusion_fraction), b_u, u, iteration_count, BiCGSTAB, 0.5, 1e-6, PC_Jacobi);
        iterative_solve(csr_blend(a_v, 1. - diffusion_fraction, a_di, diffusion_fraction), b_v, v, iteration_count, BiCGSTAB, 0.5, 1e-6, PC_Jacobi);
        iterative_solve(csr_blend(a_w, 1. - diffusion_fraction, a_di, diffusion_fraction), b_w, w, iteration_count, BiCGSTAB, 0.5, 1e-6, PC_Jacobi);
        diffusion_fraction -= 0.2;
    }
}

// The potential system of initialize_velocity_field (:524-590): grad psi = velocity. VelocityInlet faces give the source
// -(zone velocity . n_out); PressureOutlet fixes psi = 0 on the face; walls, symmetry and everything else are natural boundaries.
// Note the outlet coefficient carries no area / volume factor (:561-568), unlike the interior one: restated as written.
void build_velocity_potential(const Mesh& m, Csr& a_out, DVec& b) {
    size_t n = m.cells.size();
    Coo a; a.nrows = a.ncols = n;
    b.assign(n, 0.);
    for (size_t ci = 0; ci < n; ++ci) {
        const Cell& cell = m.cells[ci];
        Float a_p = 0.;
        for (size_t fi : cell.face_indices) {
            const Face& face = m.faces[fi];
            Vec3 nout = get_outward_face_normal(face, ci);
            const FaceZone& z = zone_of(m, face);
            Float a_nb, source; size_t nb;
            switch (z.zone_type) {
                case Interior:
                    nb = face.cell_indices[0] == ci ? face.cell_indices[1] : face.cell_indices[0];
                    a_nb = vdot(vreciprocal(cell.centroid - m.cells[nb].centroid), nout) * (face.area / cell.volume);
                    source = 0.;
                    break;
                case VelocityInlet: a_nb = 0.; source = -vdot(z.vector_value, nout); nb = NONE; break;
                case PressureOutlet: a_nb = vdot(vreciprocal(cell.centroid - face.centroid), nout); source = 0.; nb = NONE; break;
                default: a_nb = 0.; source = 0.; nb = NONE;  // Symmetry | Wall | everything else (:556-575)
            }
            if (nb != NONE) a.push(ci, nb, -a_nb);
            b[ci] += source;
            a_p += a_nb;
        }
        a.push(ci, ci, a_p);
    }
    a_out = coo_to_csr(a);
}

// The least-squares gradient of psi over the cell neighbours (:624-693). Columns of the neighbour matrix that are entirely zero
// are dropped before the normal equations (a one-cell-thick mesh has no z differences); a singular system leaves the cell at
// zero ("Could not invert. Skipping."), NaN components become zero.
Vec3 potential_gradient(const Mesh& m, const DVec& psi, size_t ci) {
    const Cell& cell = m.cells[ci];
    std::vector<Float> a_data, b_data;
    size_t neighbor_count = 0;
    for (size_t fi : cell.face_indices) {
        const Face& f = m.faces[fi];
        if (f.cell_indices.size() != 2) continue;
        size_t nb = f.cell_indices[0] != ci ? f.cell_indices[0] : f.cell_indices[1];
        Vec3 dx = m.cells[nb].centroid - cell.centroid;
        a_data.push_back(dx.x); a_data.push_back(dx.y); a_data.push_back(dx.z);
        b_data.push_back(psi[nb] - psi[ci]);
        ++neighbor_count;
    }
    Dense a = dense_from_row_slice(neighbor_count, 3, a_data);
    std::vector<size_t> nonzero_columns;
    for (size_t j = 0; j < 3; ++j) {  // col.min() != 0. || col.max() != 0. (an empty column counts as all zero)
        bool nz = false;
        for (size_t i = 0; i < neighbor_count; ++i) if (a.at(i, j) != 0.) nz = true;
        if (nz) nonzero_columns.push_back(j);
    }
    Dense sel(neighbor_count, nonzero_columns.size());
    for (size_t q = 0; q < nonzero_columns.size(); ++q) for (size_t i = 0; i < neighbor_count; ++i) sel.at(i, q) = a.at(i, nonzero_columns[q]);
    Dense at = dense_transpose(sel);
    DVec b = dense_mul_vec(at, b_data);
    Dense ata = dense_mul(at, sel);
    DVec vel(3, 0.);
    Dense inv = ata;
    if (dense_try_inverse(inv)) {
        DVec r = dense_mul_vec(inv, b);
        for (size_t q = 0; q < r.size(); ++q) vel[q] = r[q];
    }
    auto comp = [&](size_t axis) {
        for (size_t q = 0; q < nonzero_columns.size(); ++q) if (nonzero_columns[q] == axis) return vel[q];
        return Float(0.);
    };
    Vec3 g{comp(0), comp(1), comp(2)};
    if (g.x != g.x) g.x = 0.;
    if (g.y != g.y) g.y = 0.;
    if (g.z != g.z) g.z = 0.;
    return g;
}

void initialize_velocity_field(const Mesh& m, DVec& u, DVec& v, DVec& w) {  // :511-696 (its two debug files are not written)
    Csr a; DVec b;
    build_velocity_potential(m, a, b);
    DVec psi(m.cells.size(), 0.);
    iterative_solve(a, b, psi, 10, BiCGSTAB, 0.1, 1e-6, PC_Jacobi);  // :592-601
    for (size_t ci = 0; ci < m.cells.size(); ++ci) {
        Vec3 g = potential_gradient(m, psi, ci);
        u[ci] = g.x; v[ci] = g.y; w[ci] = g.z;
    }
}

// initialize_flow_new (:354-410). The match arms overlap: `PressureOnly | Hybrid` comes first, so a Hybrid system only gets its
// pressure field initialised and the velocity field stays zero.
void initialize_flow_new(const Mesh& m, Float, Float, uint64_t, DVec& u, DVec& v, DVec& w, DVec& p) {
    size_t n = m.cells.size();
    u.assign(n, 0.); v.assign(n, 0.); w.assign(n, 0.); p.assign(n, 0.);
    int constraint = check_boundary_conditions(m);
    if (constraint == PressureOnly || constraint == Hybrid) initialize_pressure_field(m, p);
    else initialize_velocity_field(m, u, v, w);
}

}  // namespace orc_oracle
