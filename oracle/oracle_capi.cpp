// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH. See orc_oracle.hpp.
// Flat C entry points over the C++ restatement so that tests/, smoke() and bench.py's cpu_baseline
// leg can drive it through ctypes. Every call returns 0 on success and 1 when the restated code
// "panicked" (message retrievable with oo_last_error).
#include <cstring>
#include <string>

#include "orc_oracle.hpp"

using namespace orc_oracle;

static thread_local std::string g_err;
#define OO_TRY(...)                       \
    try { __VA_ARGS__; return 0; }        \
    catch (const std::exception& e) { g_err = e.what(); return 1; }

static Csr csr_from(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int64_t* col, const double* val) {
    Csr a;
    a.nrows = size_t(nrows); a.ncols = size_t(ncols);
    a.rowptr.assign(rowptr, rowptr + nrows + 1);
    a.col.assign(col, col + rowptr[nrows]);
    a.val.assign(val, val + rowptr[nrows]);
    return a;
}
static NumericalSettings settings_from(const int64_t* iv, const double* dv) {
    // iv: momentum, limiter, pressure_interp, velocity_interp, gradient, solver_type, iterations, preconditioner,
    //     gs_intended, mg_smoother, mg_levels      dv: p_relax, u_relax, relaxation, threshold
    NumericalSettings s;
    s.momentum = int(iv[0]); s.limiter = int(iv[1]); s.pressure_interpolation = int(iv[2]); s.velocity_interpolation = int(iv[3]);
    s.gradient_reconstruction = int(iv[4]); s.matrix_solver.solver_type = int(iv[5]); s.matrix_solver.iterations = uint64_t(iv[6]);
    s.matrix_solver.preconditioner = int(iv[7]); s.opts.gs_intended = iv[8] != 0; s.opts.mg_smoother = int(iv[9]); s.opts.mg_levels = uint64_t(iv[10]);
    s.pressure_relaxation = dv[0]; s.momentum_relaxation = dv[1]; s.matrix_solver.relaxation = dv[2];
    s.matrix_solver.relative_convergence_threshold = dv[3];
    return s;
}
static DVec dv(const double* p, size_t n) { return DVec(p, p + n); }

extern "C" {

const char* oo_last_error() { return g_err.c_str(); }

// ---- meshes ---------------------------------------------------------------------------------------
int oo_mesh_read(const char* path, void** out) { OO_TRY(*out = new Mesh(read_mesh(path))); }
int oo_mesh_from_arrays(int dims, int64_t n_nodes, const double* xyz, int64_t n_faces, const int64_t* face_node_offsets,
                        const int64_t* face_nodes, const int64_t* c0, const int64_t* c1, const int64_t* face_zone, int64_t n_zones,
                        const int64_t* zone_ids, const int64_t* zone_types, const char* const* zone_names, void** out) {
    OO_TRY(*out = new Mesh(mesh_from_arrays(dims, size_t(n_nodes), xyz, size_t(n_faces), face_node_offsets, face_nodes, c0, c1, face_zone,
                                            size_t(n_zones), zone_ids, zone_types, zone_names)));
}
void oo_mesh_free(void* m) { delete static_cast<Mesh*>(m); }
void oo_mesh_counts(void* mp, int64_t* out) {  // cells, faces, nodes, zones, sum of cell face-list lengths, dims
    Mesh& m = *static_cast<Mesh*>(mp);
    size_t tot = 0;
    for (auto& c : m.cells) tot += c.face_indices.size();
    out[0] = int64_t(m.cells.size()); out[1] = int64_t(m.faces.size()); out[2] = int64_t(m.vertices.size());
    out[3] = int64_t(m.face_zones.size()); out[4] = int64_t(tot); out[5] = m.dimensions;
}
void oo_mesh_export(void* mp, int64_t* face_c0, int64_t* face_c1, int64_t* face_zone, double* face_area, double* face_normal,
                    double* face_centroid, double* cell_volume, double* cell_centroid, int64_t* cell_face_offsets, int64_t* cell_face_indices) {
    Mesh& m = *static_cast<Mesh*>(mp);
    for (size_t f = 0; f < m.faces.size(); ++f) {
        const Face& fa = m.faces[f];
        face_c0[f] = int64_t(fa.cell_indices[0]);
        face_c1[f] = fa.cell_indices.size() > 1 ? int64_t(fa.cell_indices[1]) : -1;
        face_zone[f] = int64_t(fa.zone);
        face_area[f] = fa.area;
        face_normal[3 * f] = fa.normal.x; face_normal[3 * f + 1] = fa.normal.y; face_normal[3 * f + 2] = fa.normal.z;
        face_centroid[3 * f] = fa.centroid.x; face_centroid[3 * f + 1] = fa.centroid.y; face_centroid[3 * f + 2] = fa.centroid.z;
    }
    size_t off = 0;
    for (size_t c = 0; c < m.cells.size(); ++c) {
        const Cell& ce = m.cells[c];
        cell_volume[c] = ce.volume;
        cell_centroid[3 * c] = ce.centroid.x; cell_centroid[3 * c + 1] = ce.centroid.y; cell_centroid[3 * c + 2] = ce.centroid.z;
        cell_face_offsets[c] = int64_t(off);
        for (size_t fi : ce.face_indices) cell_face_indices[off++] = int64_t(fi);
    }
    cell_face_offsets[m.cells.size()] = int64_t(off);
}
// zone table in ascending zone-id order: ids, types, scalar, vector[3]; names copied into 64-byte slots
void oo_mesh_zones(void* mp, int64_t* ids, int64_t* types, double* scalar, double* vec, char* names64) {
    Mesh& m = *static_cast<Mesh*>(mp);
    size_t k = 0;
    for (auto& kv : m.face_zones) {
        ids[k] = int64_t(kv.first); types[k] = kv.second.zone_type; scalar[k] = kv.second.scalar_value;
        vec[3 * k] = kv.second.vector_value.x; vec[3 * k + 1] = kv.second.vector_value.y; vec[3 * k + 2] = kv.second.vector_value.z;
        std::strncpy(names64 + 64 * k, kv.second.name.c_str(), 63);
        names64[64 * k + 63] = 0;
        ++k;
    }
}
int oo_mesh_set_zone(void* mp, const char* name, int64_t type, double scalar, double vx, double vy, double vz) {
    OO_TRY({
        FaceZone& fz = static_cast<Mesh*>(mp)->get_face_zone(name);
        fz.zone_type = int(type); fz.scalar_value = scalar; fz.vector_value = {vx, vy, vz};
    });
}

// ---- CSR handles ----------------------------------------------------------------------------------
int oo_csr_new(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int64_t* col, const double* val, void** out) {
    OO_TRY(*out = new Csr(csr_from(nrows, ncols, rowptr, col, val)));
}
void oo_csr_free(void* a) { delete static_cast<Csr*>(a); }
void oo_csr_dims(void* ap, int64_t* out) { Csr& a = *static_cast<Csr*>(ap); out[0] = int64_t(a.nrows); out[1] = int64_t(a.ncols); out[2] = int64_t(a.nnz()); }
void oo_csr_get(void* ap, int64_t* rowptr, int64_t* col, double* val) {
    Csr& a = *static_cast<Csr*>(ap);
    for (size_t i = 0; i <= a.nrows; ++i) rowptr[i] = int64_t(a.rowptr[i]);
    for (size_t k = 0; k < a.nnz(); ++k) { col[k] = int64_t(a.col[k]); val[k] = a.val[k]; }
}
void oo_csr_set_values(void* ap, const double* val) { Csr& a = *static_cast<Csr*>(ap); a.val.assign(val, val + a.nnz()); }
int oo_spmv(void* ap, const double* x, double* y) {
    OO_TRY({ Csr& a = *static_cast<Csr*>(ap); DVec r = spmv(a, dv(x, a.ncols)); std::memcpy(y, r.data(), 8 * r.size()); });
}
int oo_spgemm(void* a, void* b, void** out) { OO_TRY(*out = new Csr(spgemm(*static_cast<Csr*>(a), *static_cast<Csr*>(b)))); }
int oo_transpose(void* a, void** out) { OO_TRY(*out = new Csr(transpose(*static_cast<Csr*>(a)))); }
int oo_jacobi_scale(void* ap, const double* b, void** a_out, double* b_out) {  // linear_algebra.rs:159-167
    OO_TRY({
        Csr& a = *static_cast<Csr*>(ap);
        Csr p_inv = diagonal_as_csr(a);
        for (double& v : p_inv.val) v = 1. / v;
        *a_out = new Csr(spgemm(p_inv, a));
        DVec r = spmv(p_inv, dv(b, a.nrows));
        std::memcpy(b_out, r.data(), 8 * r.size());
    });
}
double oo_dot(const double* a, const double* b, int64_t n) { return dot(dv(a, size_t(n)), dv(b, size_t(n))); }
int oo_build_restriction(void* a, int64_t method, void** out) { OO_TRY(*out = new Csr(build_restriction_matrix(*static_cast<Csr*>(a), int(method)))); }
int oo_galerkin(void* r, void* a, void** out) {  // linear_algebra.rs:84
    OO_TRY({ Csr& R = *static_cast<Csr*>(r); *out = new Csr(spgemm(spgemm(R, *static_cast<Csr*>(a)), transpose(R))); });
}
int oo_iterative_solve(void* ap, const double* b, double* x, int64_t iterations, int64_t method, double relaxation, double threshold,
                       int64_t preconditioner, int64_t gs_intended, int64_t mg_smoother, int64_t mg_levels) {
    OO_TRY({
        Csr& a = *static_cast<Csr*>(ap);
        DVec xv = dv(x, a.ncols);
        SolveOpts o; o.gs_intended = gs_intended != 0; o.mg_smoother = int(mg_smoother); o.mg_levels = uint64_t(mg_levels);
        iterative_solve(a, dv(b, a.nrows), xv, uint64_t(iterations), int(method), relaxation, threshold, int(preconditioner), o);
        std::memcpy(x, xv.data(), 8 * xv.size());
    });
}
// Multigrid trace: run a Multigrid solve and keep R_l, A_l of every level; returns the level count.
static MgTrace g_mg;
int oo_mg_trace_solve(void* ap, const double* b, double* x, int64_t iterations, double relaxation, double threshold, int64_t preconditioner,
                      int64_t mg_smoother, int64_t mg_levels, int64_t* n_levels) {
    g_mg = MgTrace();
    set_mg_trace(&g_mg);
    int rc = oo_iterative_solve(ap, b, x, iterations, Multigrid, relaxation, threshold, preconditioner, 1, mg_smoother, mg_levels);
    set_mg_trace(nullptr);
    *n_levels = int64_t(g_mg.restriction.size());
    return rc;
}
void* oo_mg_trace_restriction(int64_t l) { return new Csr(g_mg.restriction[size_t(l)]); }
void* oo_mg_trace_coarse(int64_t l) { return new Csr(g_mg.coarse[size_t(l)]); }

// ---- discretization / solver ------------------------------------------------------------------------
int oo_build_momentum_diffusion(void* m, double mu, void** a_out, double* b_u, double* b_v, double* b_w) {
    OO_TRY({
        Csr a; DVec bu, bv, bw;
        build_momentum_diffusion_matrix(*static_cast<Mesh*>(m), mu, a, bu, bv, bw);
        *a_out = new Csr(a);
        std::memcpy(b_u, bu.data(), 8 * bu.size()); std::memcpy(b_v, bv.data(), 8 * bv.size()); std::memcpy(b_w, bw.data(), 8 * bw.size());
    });
}
int oo_init_momentum_matrix(void* m, void** out) { OO_TRY(*out = new Csr(initialize_momentum_matrix(*static_cast<Mesh*>(m)))); }
int oo_build_momentum_advection(void* mp, void* a_u, void* a_v, void* a_w, void* a_di, const double* u, const double* v, const double* w,
                                const double* p, const int64_t* iv, const double* dvv, double rho, double* b_u, double* b_v, double* b_w,
                                double* peclet3) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        NumericalSettings s = settings_from(iv, dvv);
        DVec bu(n, 0.), bv(n, 0.), bw(n, 0.);
        Peclet pe = build_momentum_advection_matrices(*static_cast<Csr*>(a_u), *static_cast<Csr*>(a_v), *static_cast<Csr*>(a_w), bu, bv, bw,
                                                      *static_cast<Csr*>(a_di), m, dv(u, n), dv(v, n), dv(w, n), dv(p, n), s.momentum, s.limiter,
                                                      s.velocity_interpolation, s.pressure_interpolation, s.gradient_reconstruction, rho);
        std::memcpy(b_u, bu.data(), 8 * n); std::memcpy(b_v, bv.data(), 8 * n); std::memcpy(b_w, bw.data(), 8 * n);
        peclet3[0] = pe.avg; peclet3[1] = pe.min; peclet3[2] = pe.max;
    });
}
int oo_build_pressure_correction(void* mp, void* a_u, void* a_v, void* a_w, const double* u, const double* v, const double* w, const double* p,
                                 const int64_t* iv, const double* dvv, double rho, void** a_out, double* b_out) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        Csr a; DVec b;
        build_pressure_correction_matrices(m, dv(u, n), dv(v, n), dv(w, n), dv(p, n), *static_cast<Csr*>(a_u), *static_cast<Csr*>(a_v),
                                           *static_cast<Csr*>(a_w), settings_from(iv, dvv), rho, a, b);
        *a_out = new Csr(a);
        std::memcpy(b_out, b.data(), 8 * n);
    });
}
int oo_apply_pressure_correction(void* mp, void* a_u, void* a_v, void* a_w, const double* p_prime, double* u, double* v, double* w, double* p,
                                 const int64_t* iv, const double* dvv, double* norms2) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec uu = dv(u, n), vv = dv(v, n), ww = dv(w, n), pp = dv(p, n);
        CorrectionNorms cn = apply_pressure_correction(m, *static_cast<Csr*>(a_u), *static_cast<Csr*>(a_v), *static_cast<Csr*>(a_w), dv(p_prime, n),
                                                       uu, vv, ww, pp, settings_from(iv, dvv));
        std::memcpy(u, uu.data(), 8 * n); std::memcpy(v, vv.data(), 8 * n); std::memcpy(w, ww.data(), 8 * n); std::memcpy(p, pp.data(), 8 * n);
        norms2[0] = cn.p_prime_norm; norms2[1] = cn.velocity_corr;
    });
}
int oo_pressure_gradient(void* mp, const double* p, double* grad3n) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec pp = dv(p, n);
        for (size_t c = 0; c < n; ++c) { Vec3 g = calculate_pressure_gradient(m, pp, c, G_GreenGaussCell); grad3n[3 * c] = g.x; grad3n[3 * c + 1] = g.y; grad3n[3 * c + 2] = g.z; }
    });
}
struct ReportSink { double* out; int64_t cap, n; };
static void report_cb(const IterationReport* r, void* user) {
    ReportSink* s = static_cast<ReportSink*>(user);
    if (s->n >= s->cap) return;
    double* o = s->out + 9 * s->n++;
    o[0] = double(r->iteration); o[1] = r->u_avg; o[2] = r->v_avg; o[3] = r->w_avg; o[4] = r->peclet_avg; o[5] = r->peclet_min;
    o[6] = r->peclet_max; o[7] = r->vel_corr; o[8] = r->p_corr;
}
int oo_solve_steady(void* mp, double* u, double* v, double* w, double* p, const int64_t* iv, const double* dvv, double rho, double mu,
                    int64_t iterations, int64_t report_every, double* reports9, int64_t reports_cap, int64_t* n_reports, double* phase_times5) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec uu = dv(u, n), vv = dv(v, n), ww = dv(w, n), pp = dv(p, n);
        ReportSink sink{reports9, reports_cap, 0};
        PhaseTimes pt;
        try {
            solve_steady(m, uu, vv, ww, pp, settings_from(iv, dvv), rho, mu, uint64_t(iterations), uint64_t(report_every),
                         reports9 ? report_cb : nullptr, &sink, &pt);
        } catch (...) {
            if (n_reports) *n_reports = sink.n;
            throw;
        }
        if (n_reports) *n_reports = sink.n;
        if (phase_times5) { phase_times5[0] = pt.momentum_assembly; phase_times5[1] = pt.momentum_solves; phase_times5[2] = pt.pressure_assembly; phase_times5[3] = pt.pressure_solve; phase_times5[4] = pt.correction; }
        std::memcpy(u, uu.data(), 8 * n); std::memcpy(v, vv.data(), 8 * n); std::memcpy(w, ww.data(), 8 * n); std::memcpy(p, pp.data(), 8 * n);
    });
}

int oo_check_boundary_conditions(void* mp, int64_t* type_out) { OO_TRY(*type_out = check_boundary_conditions(*static_cast<Mesh*>(mp))); }
int oo_build_pressure_laplace(void* mp, void** a_out, double* b_out) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        Csr a; DVec b;
        build_pressure_laplace(m, a, b);
        *a_out = new Csr(a);
        std::memcpy(b_out, b.data(), 8 * b.size());
    });
}
int oo_initialize_flow(void* mp, double mu, double rho, int64_t iteration_count, double* u, double* v, double* w, double* p) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec uu, vv, ww, pp;
        initialize_flow(m, mu, rho, uint64_t(iteration_count), uu, vv, ww, pp);
        std::memcpy(u, uu.data(), 8 * n); std::memcpy(v, vv.data(), 8 * n); std::memcpy(w, ww.data(), 8 * n); std::memcpy(p, pp.data(), 8 * n);
    });
}

int oo_gradients(void* mp, const double* u, const double* v, const double* w, const double* p, int64_t gradient, double* grad_p3n, double* grad_u9n) {
    OO_TRY({  // calculate_pressure_gradient / calculate_velocity_gradient for every cell (what write_gradients prints, src/io.rs:623-662)
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec uu = dv(u, n), vv = dv(v, n), ww = dv(w, n), pp = dv(p, n);
        for (size_t c = 0; c < n; ++c) {
            if (grad_p3n) { Vec3 g = calculate_pressure_gradient(m, pp, c, int(gradient)); grad_p3n[3 * c] = g.x; grad_p3n[3 * c + 1] = g.y; grad_p3n[3 * c + 2] = g.z; }
            if (grad_u9n) {
                Tensor3 t = calculate_velocity_gradient(m, uu, vv, ww, c, int(gradient));
                double* o = grad_u9n + 9 * c;
                o[0] = t.x.x; o[1] = t.x.y; o[2] = t.x.z; o[3] = t.y.x; o[4] = t.y.y; o[5] = t.y.z; o[6] = t.z.x; o[7] = t.z.y; o[8] = t.z.z;
            }
        }
    });
}
int oo_build_velocity_potential(void* mp, void** a_out, double* b_out) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        Csr a; DVec b;
        build_velocity_potential(m, a, b);
        *a_out = new Csr(a);
        std::memcpy(b_out, b.data(), 8 * b.size());
    });
}
int oo_potential_gradient(void* mp, const double* psi, double* grad3n) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec ps = dv(psi, n);
        for (size_t c = 0; c < n; ++c) { Vec3 g = potential_gradient(m, ps, c); grad3n[3 * c] = g.x; grad3n[3 * c + 1] = g.y; grad3n[3 * c + 2] = g.z; }
    });
}
int oo_initialize_flow_new(void* mp, double mu, double rho, int64_t iteration_count, double* u, double* v, double* w, double* p) {
    OO_TRY({
        Mesh& m = *static_cast<Mesh*>(mp);
        size_t n = m.cells.size();
        DVec uu, vv, ww, pp;
        initialize_flow_new(m, mu, rho, uint64_t(iteration_count), uu, vv, ww, pp);
        std::memcpy(u, uu.data(), 8 * n); std::memcpy(v, vv.data(), 8 * n); std::memcpy(w, ww.data(), 8 * n); std::memcpy(p, pp.data(), 8 * n);
    });
}

int oo_set_partition(const int64_t* cuts, int64_t n_cuts) {   // n_cuts == 0 switches the emulation off
    OO_TRY({
        if (!cuts || n_cuts < 2) { set_partition(nullptr); }
        else { std::vector<size_t> c(cuts, cuts + n_cuts); set_partition(&c); }
    });
}

}  // extern "C"
