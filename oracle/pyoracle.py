"""ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.

ctypes driver for oracle/_build/liborc_oracle.so (the C++ CPU restatement of ORC's SIMPLE inner
loop, see orc_oracle.hpp). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker / reported baseline.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc_oracle.so")

# enum values (orc_oracle.hpp)
UD, CD1, CD2, TVD = 0, 1, 2, 3
PSI_UD, PSI_CD1, PSI_LUD, PSI_QUICK, PSI_UMIST = 0, 1, 2, 3, 4
P_LINEAR, P_LINEAR_WEIGHTED, P_STANDARD, P_SECOND_ORDER = 0, 1, 2, 3
V_LINEAR, V_LINEAR_WEIGHTED, V_RHIE_CHOW = 0, 1, 2
GAUSS_SEIDEL, JACOBI, MULTIGRID, BICGSTAB = 0, 1, 2, 3
PC_NONE, PC_JACOBI = 0, 1
INJECTION, STRONGEST = 0, 1
INTERIOR, WALL, PRESSURE_INLET, PRESSURE_OUTLET, SYMMETRY, VELOCITY_INLET = 2, 3, 4, 5, 7, 10


class OraclePanic(RuntimeError):
    """The restated code hit one of the reference's panic!() sites."""


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("orc_oracle.cpp", "oracle_capi.cpp", "orc_oracle.hpp")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oo_last_error.restype = C.c_char_p
        _lib.oo_dot.restype = C.c_double
        _lib.oo_mg_trace_restriction.restype = C.c_void_p
        _lib.oo_mg_trace_coarse.restype = C.c_void_p
    return _lib


def _chk(rc):
    if rc != 0:
        raise OraclePanic(lib().oo_last_error().decode())


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Settings:
    """Mirror of NumericalSettings + MatrixSolverSettings defaults (src/lib.rs:58-86)."""

    def __init__(self, **kw):
        self.momentum = CD1
        self.limiter = PSI_QUICK
        self.pressure_interpolation = P_SECOND_ORDER
        self.velocity_interpolation = V_RHIE_CHOW
        self.gradient = 0
        self.solver_type = MULTIGRID
        self.iterations = 50
        self.preconditioner = PC_JACOBI
        self.gs_intended = 1
        self.mg_smoother = BICGSTAB
        self.mg_levels = 3
        self.pressure_relaxation = 0.01
        self.momentum_relaxation = 0.5
        self.relaxation = 0.5
        self.threshold = 1e-3
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    def pack(self):
        iv = _i64([self.momentum, self.limiter, self.pressure_interpolation, self.velocity_interpolation, self.gradient,
                   self.solver_type, self.iterations, self.preconditioner, self.gs_intended, self.mg_smoother, self.mg_levels])
        dv = _f64([self.pressure_relaxation, self.momentum_relaxation, self.relaxation, self.threshold])
        return iv, dv


class Csr:
    def __init__(self, handle):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle

    @classmethod
    def from_arrays(cls, nrows, ncols, rowptr, col, val):
        out = C.c_void_p()
        rp, co, va = _i64(rowptr), _i64(col), _f64(val)
        _chk(lib().oo_csr_new(C.c_int64(nrows), C.c_int64(ncols), _p(rp), _p(co), _p(va), C.byref(out)))
        return cls(out)

    def __del__(self):
        try:
            lib().oo_csr_free(self.h)
        except Exception:
            pass

    @property
    def dims(self):
        d = np.zeros(3, np.int64)
        lib().oo_csr_dims(self.h, _p(d))
        return int(d[0]), int(d[1]), int(d[2])

    def arrays(self):
        nr, nc, nnz = self.dims
        rp, co, va = np.zeros(nr + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz, np.float64)
        lib().oo_csr_get(self.h, _p(rp), _p(co), _p(va))
        return rp, co, va

    def set_values(self, val):
        va = _f64(val)
        assert va.size == self.dims[2]
        lib().oo_csr_set_values(self.h, _p(va))

    def spmv(self, x):
        nr, nc, _ = self.dims
        x = _f64(x)
        assert x.size == nc
        y = np.zeros(nr)
        _chk(lib().oo_spmv(self.h, _p(x), _p(y)))
        return y

    def matmul(self, other):
        out = C.c_void_p()
        _chk(lib().oo_spgemm(self.h, other.h, C.byref(out)))
        return Csr(out)

    def transpose(self):
        out = C.c_void_p()
        _chk(lib().oo_transpose(self.h, C.byref(out)))
        return Csr(out)

    def jacobi_scale(self, b):
        b = _f64(b)
        out = C.c_void_p()
        bo = np.zeros_like(b)
        _chk(lib().oo_jacobi_scale(self.h, _p(b), C.byref(out), _p(bo)))
        return Csr(out), bo

    def build_restriction(self, method=STRONGEST):
        out = C.c_void_p()
        _chk(lib().oo_build_restriction(self.h, C.c_int64(method), C.byref(out)))
        return Csr(out)

    def to_scipy(self):
        import scipy.sparse as sp
        rp, co, va = self.arrays()
        nr, nc, _ = self.dims
        return sp.csr_matrix((va, co, rp), shape=(nr, nc))


def galerkin(r, a):
    out = C.c_void_p()
    _chk(lib().oo_galerkin(r.h, a.h, C.byref(out)))
    return Csr(out)


def dot(a, b):
    a, b = _f64(a), _f64(b)
    return float(lib().oo_dot(_p(a), _p(b), C.c_int64(a.size)))


def iterative_solve(a, b, x, iterations, method, relaxation, threshold, preconditioner, gs_intended=1, mg_smoother=BICGSTAB, mg_levels=3):
    """src/linear_algebra.rs:144-299. Returns the updated solution vector (x is not modified)."""
    b, x = _f64(b), _f64(x).copy()
    _chk(lib().oo_iterative_solve(a.h, _p(b), _p(x), C.c_int64(iterations), C.c_int64(method), C.c_double(relaxation), C.c_double(threshold),
                                  C.c_int64(preconditioner), C.c_int64(gs_intended), C.c_int64(mg_smoother), C.c_int64(mg_levels)))
    return x


def multigrid_trace(a, b, x, iterations=50, relaxation=0.5, threshold=1e-3, preconditioner=PC_JACOBI, mg_smoother=BICGSTAB, mg_levels=3):
    """Multigrid solve that also returns [(R_l, A_l)] for every coarse level built (linear_algebra.rs:80-84)."""
    b, x = _f64(b), _f64(x).copy()
    nl = C.c_int64()
    _chk(lib().oo_mg_trace_solve(a.h, _p(b), _p(x), C.c_int64(iterations), C.c_double(relaxation), C.c_double(threshold), C.c_int64(preconditioner),
                                 C.c_int64(mg_smoother), C.c_int64(mg_levels), C.byref(nl)))
    levels = [(Csr(lib().oo_mg_trace_restriction(C.c_int64(l))), Csr(lib().oo_mg_trace_coarse(C.c_int64(l)))) for l in range(nl.value)]
    return x, levels


def set_partition(cuts=None):
    """Partition emulation (not in the reference): with cuts [0, c_1, ..., N] the oracle restates the multi-GPU path's two documented
    deviations (partition-lagged diagonals in the momentum assembly, Multigrid coarse correction per partition block). None = off."""
    if cuts is None:
        _chk(lib().oo_set_partition(None, C.c_int64(0)))
    else:
        c = _i64(cuts)
        _chk(lib().oo_set_partition(_p(c), C.c_int64(c.size)))


class Mesh:
    def __init__(self, handle):
        self.h = handle

    @classmethod
    def read(cls, path):
        out = C.c_void_p()
        _chk(lib().oo_mesh_read(path.encode(), C.byref(out)))
        return cls(out)

    @classmethod
    def from_arrays(cls, dims, xyz, face_node_offsets, face_nodes, c0, c1, face_zone, zone_ids, zone_types, zone_names):
        out = C.c_void_p()
        xyz = _f64(xyz)
        fo, fn, a0, a1, fz = _i64(face_node_offsets), _i64(face_nodes), _i64(c0), _i64(c1), _i64(face_zone)
        zi, zt = _i64(zone_ids), _i64(zone_types)
        names = (C.c_char_p * len(zone_names))(*[n.encode() for n in zone_names])
        _chk(lib().oo_mesh_from_arrays(C.c_int(dims), C.c_int64(xyz.size // 3), _p(xyz), C.c_int64(a0.size), _p(fo), _p(fn), _p(a0), _p(a1), _p(fz),
                                       C.c_int64(zi.size), _p(zi), _p(zt), names, C.byref(out)))
        return cls(out)

    def __del__(self):
        try:
            lib().oo_mesh_free(self.h)
        except Exception:
            pass

    def counts(self):
        d = np.zeros(6, np.int64)
        lib().oo_mesh_counts(self.h, _p(d))
        return dict(cells=int(d[0]), faces=int(d[1]), nodes=int(d[2]), zones=int(d[3]), cell_faces=int(d[4]), dims=int(d[5]))

    @property
    def n_cells(self):
        return self.counts()["cells"]

    def export(self):
        c = self.counts()
        nf, nc = c["faces"], c["cells"]
        out = dict(face_c0=np.zeros(nf, np.int64), face_c1=np.zeros(nf, np.int64), face_zone=np.zeros(nf, np.int64), face_area=np.zeros(nf),
                   face_normal=np.zeros((nf, 3)), face_centroid=np.zeros((nf, 3)), cell_volume=np.zeros(nc), cell_centroid=np.zeros((nc, 3)),
                   cell_face_offsets=np.zeros(nc + 1, np.int64), cell_face_indices=np.zeros(c["cell_faces"], np.int64))
        lib().oo_mesh_export(self.h, *[_p(out[k]) for k in ("face_c0", "face_c1", "face_zone", "face_area", "face_normal", "face_centroid",
                                                             "cell_volume", "cell_centroid", "cell_face_offsets", "cell_face_indices")])
        return out

    def zones(self):
        nz = self.counts()["zones"]
        ids, types, sc, vec = np.zeros(nz, np.int64), np.zeros(nz, np.int64), np.zeros(nz), np.zeros((nz, 3))
        names = C.create_string_buffer(64 * nz)
        lib().oo_mesh_zones(self.h, _p(ids), _p(types), _p(sc), _p(vec), names)
        nm = [names.raw[64 * k:64 * (k + 1)].split(b"\0")[0].decode() for k in range(nz)]
        return dict(ids=ids, types=types, scalar=sc, vector=vec, names=nm)

    def set_zone(self, name, zone_type, scalar=0.0, vector=(0.0, 0.0, 0.0)):
        _chk(lib().oo_mesh_set_zone(self.h, name.encode(), C.c_int64(zone_type), C.c_double(scalar), C.c_double(vector[0]), C.c_double(vector[1]),
                                    C.c_double(vector[2])))

    # ---- discretization.rs ----
    def build_momentum_diffusion(self, mu):
        n = self.n_cells
        out = C.c_void_p()
        bu, bv, bw = np.zeros(n), np.zeros(n), np.zeros(n)
        _chk(lib().oo_build_momentum_diffusion(self.h, C.c_double(mu), C.byref(out), _p(bu), _p(bv), _p(bw)))
        return Csr(out), bu, bv, bw

    def init_momentum_matrix(self):
        out = C.c_void_p()
        _chk(lib().oo_init_momentum_matrix(self.h, C.byref(out)))
        return Csr(out)

    def build_momentum_advection(self, a_u, a_v, a_w, a_di, u, v, w, p, settings, rho):
        """In place on a_u/a_v/a_w (like the reference); returns (b_u, b_v, b_w, (pe_avg, pe_min, pe_max))."""
        n = self.n_cells
        iv, dv = settings.pack()
        u, v, w, p = _f64(u), _f64(v), _f64(w), _f64(p)
        bu, bv, bw, pe = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(3)
        _chk(lib().oo_build_momentum_advection(self.h, a_u.h, a_v.h, a_w.h, a_di.h, _p(u), _p(v), _p(w), _p(p), _p(iv), _p(dv), C.c_double(rho),
                                               _p(bu), _p(bv), _p(bw), _p(pe)))
        return bu, bv, bw, tuple(pe)

    def build_pressure_correction(self, a_u, a_v, a_w, u, v, w, p, settings, rho):
        n = self.n_cells
        iv, dv = settings.pack()
        u, v, w, p = _f64(u), _f64(v), _f64(w), _f64(p)
        out = C.c_void_p()
        b = np.zeros(n)
        _chk(lib().oo_build_pressure_correction(self.h, a_u.h, a_v.h, a_w.h, _p(u), _p(v), _p(w), _p(p), _p(iv), _p(dv), C.c_double(rho),
                                                C.byref(out), _p(b)))
        return Csr(out), b

    def apply_pressure_correction(self, a_u, a_v, a_w, p_prime, u, v, w, p, settings):
        iv, dv = settings.pack()
        pp = _f64(p_prime)
        u, v, w, p = _f64(u).copy(), _f64(v).copy(), _f64(w).copy(), _f64(p).copy()
        norms = np.zeros(2)
        _chk(lib().oo_apply_pressure_correction(self.h, a_u.h, a_v.h, a_w.h, _p(pp), _p(u), _p(v), _p(w), _p(p), _p(iv), _p(dv), _p(norms)))
        return u, v, w, p, tuple(norms)

    def pressure_gradient(self, p):
        p = _f64(p)
        g = np.zeros((self.n_cells, 3))
        _chk(lib().oo_pressure_gradient(self.h, _p(p), _p(g)))
        return g

    def solve_steady(self, u, v, w, p, settings, rho, mu, iterations, report_every=1):
        """src/solver.rs:26-244. Returns (u, v, w, p, reports[k,9], phase_times[5])."""
        iv, dv = settings.pack()
        u, v, w, p = _f64(u).copy(), _f64(v).copy(), _f64(w).copy(), _f64(p).copy()
        cap = max(1, iterations // max(1, report_every))
        reports = np.zeros((cap, 9))
        nrep = C.c_int64()
        times = np.zeros(5)
        _chk(lib().oo_solve_steady(self.h, _p(u), _p(v), _p(w), _p(p), _p(iv), _p(dv), C.c_double(rho), C.c_double(mu), C.c_int64(iterations),
                                   C.c_int64(report_every), _p(reports), C.c_int64(cap), C.byref(nrep), _p(times)))
        return u, v, w, p, reports[:nrep.value], times

    # ---- flow initialisation (src/solver.rs:246-352, 414-509, 710-770) ----
    def check_boundary_conditions(self):
        t = C.c_int64()
        _chk(lib().oo_check_boundary_conditions(self.h, C.byref(t)))
        return int(t.value)   # 0 PressureOnly, 1 VelocityOnly, 2 Hybrid

    def build_pressure_laplace(self):
        n = self.n_cells
        out = C.c_void_p()
        b = np.zeros(n)
        _chk(lib().oo_build_pressure_laplace(self.h, C.byref(out), _p(b)))
        return Csr(out), b

    def initialize_flow(self, mu, rho, iteration_count):
        n = self.n_cells
        u, v, w, p = (np.zeros(n) for _ in range(4))
        _chk(lib().oo_initialize_flow(self.h, C.c_double(mu), C.c_double(rho), C.c_int64(iteration_count), _p(u), _p(v), _p(w), _p(p)))
        return u, v, w, p

    def initialize_flow_new(self, mu, rho, iteration_count):
        """src/solver.rs:354-410 -> (u, v, w, p)."""
        n = self.n_cells
        u, v, w, p = (np.zeros(n) for _ in range(4))
        _chk(lib().oo_initialize_flow_new(self.h, C.c_double(mu), C.c_double(rho), C.c_int64(iteration_count), _p(u), _p(v), _p(w), _p(p)))
        return u, v, w, p

    def build_velocity_potential(self):
        """The psi system of initialize_velocity_field (src/solver.rs:524-590) -> (a, b)."""
        n = self.n_cells
        out = C.c_void_p()
        b = np.zeros(n)
        _chk(lib().oo_build_velocity_potential(self.h, C.byref(out), _p(b)))
        return Csr(out), b

    def potential_gradient(self, psi):
        """Least-squares gradient of psi over the cell neighbours (src/solver.rs:624-693) -> (N, 3)."""
        psi = _f64(psi)
        g = np.zeros((self.n_cells, 3))
        _chk(lib().oo_potential_gradient(self.h, _p(psi), _p(g)))
        return g

    def gradients(self, u, v, w, p, gradient=0):
        """calculate_pressure_gradient / calculate_velocity_gradient of every cell (src/solver.rs:774-949) -> ((N, 3), (N, 3, 3)).
        gradient: 0 Green-Gauss cell based, 2 least squares."""
        n = self.n_cells
        u, v, w, p = _f64(u), _f64(v), _f64(w), _f64(p)
        gp, gu = np.zeros((n, 3)), np.zeros((n, 3, 3))
        _chk(lib().oo_gradients(self.h, _p(u), _p(v), _p(w), _p(p), C.c_int64(gradient), _p(gp), _p(gu)))
        return gp, gu
