"""Mirror of the reference's linear_algebra module (src/linear_algebra.rs): iterative_solve and the AMG pieces."""
import ctypes as C

import numpy as np

from . import _lib
from .context import default_context
from .mesh import _f64, _i64, _p
from .settings import NumericalSettings, MatrixSolverSettings, SolutionMethod, PreconditionMethod, RestrictionMethods


class CsrMatrix:
    """Device-resident nalgebra-sparse CsrMatrix<f64> (orc_csr handle)."""

    def __init__(self, handle, ctx):
        self._h = handle
        self.ctx = ctx

    @classmethod
    def from_arrays(cls, nrows, ncols, rowptr, col, val, ctx=None):
        ctx = ctx or default_context()
        out = C.c_void_p()
        rp, co, va = _i64(rowptr), _i64(col), _f64(val)
        _lib.check(_lib.lib().orc_csr_upload(ctx.handle, C.c_int64(nrows), C.c_int64(ncols), _p(rp), _p(co), _p(va), C.byref(out)))
        return cls(out, ctx)

    @classmethod
    def from_scipy(cls, a, ctx=None):
        a = a.tocsr()
        a.sort_indices()
        return cls.from_arrays(a.shape[0], a.shape[1], a.indptr, a.indices, a.data, ctx)

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h:
                _lib.lib().orc_csr_free(self.ctx.handle, self._h)
                self._h = None
        except Exception:
            pass

    @property
    def dims(self):
        d = np.zeros(3, np.int64)
        _lib.check(_lib.lib().orc_csr_dims(self._h, _p(d)))
        return int(d[0]), int(d[1]), int(d[2])

    def arrays(self):
        nr, nc, nnz = self.dims
        rp, co, va = np.zeros(nr + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
        _lib.check(_lib.lib().orc_csr_download(self.ctx.handle, self._h, _p(rp), _p(co), _p(va)))
        return rp, co, va

    def set_values(self, val):
        va = _f64(val)
        assert va.size == self.dims[2]
        _lib.check(_lib.lib().orc_csr_set_values(self.ctx.handle, self._h, _p(va)))

    def to_scipy(self):
        import scipy.sparse as sp
        rp, co, va = self.arrays()
        nr, nc, _ = self.dims
        return sp.csr_matrix((va, co, rp), shape=(nr, nc))

    def spmv(self, x):  # &CsrMatrix * &DVector
        nr, nc, _ = self.dims
        x = _f64(x)
        assert x.size == nc
        y = np.zeros(nr)
        _lib.check(_lib.lib().orc_spmv(self.ctx.handle, self._h, _p(x), _p(y)))
        return y

    def jacobi_scale(self, b):  # src/linear_algebra.rs:157-168
        b = _f64(b)
        out = C.c_void_p()
        bo = np.zeros_like(b)
        _lib.check(_lib.lib().orc_jacobi_scale(self.ctx.handle, self._h, _p(b), C.byref(out), _p(bo)))
        return CsrMatrix(out, self.ctx), bo


def _settings_for(iteration_count, method, relaxation_factor, convergence_threshold, preconditioner, **kw):
    from .settings import ReductionMode
    kw.setdefault("reduction_mode", ReductionMode.Fast)   # the fine-grained entries test the throughput kernels unless told otherwise
    ms = MatrixSolverSettings(solver_type=method, iterations=iteration_count, relaxation=relaxation_factor,
                              relative_convergence_threshold=convergence_threshold, preconditioner=preconditioner)
    return NumericalSettings(matrix_solver=ms, **kw).to_c()


def iterative_solve(a, b, solution_vector, iteration_count, method, relaxation_factor, convergence_threshold, preconditioner, **kw):
    """src/linear_algebra.rs:144-153, same argument order. `solution_vector` (numpy) is updated in place, like the
    reference's &mut DVector. Extra keywords: mg_smoother, mg_levels, gs_mode (constants in the reference)."""
    b = _f64(b)
    x = _f64(solution_vector).copy()
    s = _settings_for(iteration_count, method, relaxation_factor, convergence_threshold, preconditioner, **kw)
    _lib.check(_lib.lib().orc_iterative_solve(a.ctx.handle, a.handle, _p(b), _p(x), C.byref(s)))
    solution_vector[...] = x
    return solution_vector


def iterative_solve3(a, bs, xs, iteration_count, method, relaxation_factor, convergence_threshold, preconditioner, **kw):
    """Three iterative_solve calls that share `a` (the u, v, w momentum solves, src/solver.rs:99-136) in lockstep: one pass
    over the matrix per SpMV, one AMG hierarchy. `bs`, `xs`: three vectors each; `xs` are updated in place."""
    b = [_f64(q) for q in bs]
    x = [_f64(q).copy() for q in xs]
    s = _settings_for(iteration_count, method, relaxation_factor, convergence_threshold, preconditioner, **kw)
    _lib.check(_lib.lib().orc_iterative_solve3(a.ctx.handle, a.handle, _p(b[0]), _p(b[1]), _p(b[2]), _p(x[0]), _p(x[1]), _p(x[2]),
                                               C.byref(s)))
    for dst, src in zip(xs, x):
        dst[...] = src
    return xs


def build_restriction_matrix(a, method=RestrictionMethods.Strongest):  # src/linear_algebra.rs:12-63
    out = C.c_void_p()
    _lib.check(_lib.lib().orc_build_restriction(a.ctx.handle, a.handle, C.c_int32(int(method)), C.byref(out)))
    return CsrMatrix(out, a.ctx)


def galerkin(r, a):  # a' = R * a * R^T, src/linear_algebra.rs:84
    out = C.c_void_p()
    _lib.check(_lib.lib().orc_galerkin(a.ctx.handle, r.handle, a.handle, C.byref(out)))
    return CsrMatrix(out, a.ctx)


def multigrid_trace(a, b, x, iteration_count=50, relaxation_factor=0.5, convergence_threshold=1e-3,
                    preconditioner=PreconditionMethod.Jacobi, **kw):
    """Multigrid solve that also returns [(R_l, A_l)] of every coarse level (parity tests)."""
    b, x = _f64(b), _f64(x).copy()
    s = _settings_for(iteration_count, SolutionMethod.Multigrid, relaxation_factor, convergence_threshold, preconditioner, **kw)
    cap = 8
    rl, al = (C.c_void_p * cap)(), (C.c_void_p * cap)()
    n = C.c_int32()
    _lib.check(_lib.lib().orc_multigrid_trace(a.ctx.handle, a.handle, _p(b), _p(x), C.byref(s), C.c_int32(cap), rl, al, C.byref(n)))
    levels = [(CsrMatrix(C.c_void_p(rl[l]), a.ctx), CsrMatrix(C.c_void_p(al[l]), a.ctx)) for l in range(n.value)]
    return x, levels


def bench_spmv(a, reps=20, systems=1):
    ms = C.c_double()
    _lib.check(_lib.lib().orc_bench_spmv_batch(a.ctx.handle, a.handle, C.c_int32(systems), C.c_int32(reps), C.byref(ms)))
    return ms.value


def bench_bicgstab(a, reps=20, systems=1):
    ms = C.c_double()
    _lib.check(_lib.lib().orc_bench_bicgstab_batch(a.ctx.handle, a.handle, C.c_int32(systems), C.c_int32(reps), C.byref(ms)))
    return ms.value


def bench_amg_setup(a, reps=3):
    """(ms build_restriction, ms galerkin, coarse matrix) of one AMG level on `a`, device time."""
    tr, tg = C.c_double(), C.c_double()
    out = C.c_void_p()
    _lib.check(_lib.lib().orc_bench_amg_setup(a.ctx.handle, a.handle, C.c_int32(reps), C.byref(tr), C.byref(tg), C.byref(out)))
    return tr.value, tg.value, CsrMatrix(out, a.ctx)
