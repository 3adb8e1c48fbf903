"""DRY-RUN STAND-IN for libmft_b200.so (TEST INFRASTRUCTURE, opt-in, never loaded by the product or by a default pytest run).

Purpose: this container has no GPU, so `-m gpu` tests written here cannot run before the round-end hardware pass.  To keep
Python-level mistakes (argument plumbing, shapes, array layouts, oracle-side conditioning of a test case) from masking the
one thing those tests are for -- the CUDA code -- their Python halves can be executed against this stand-in:

    PYTHONPATH=tests/dryrun python -m pytest -p mockplugin tests/test_zz_* -m gpu -p no:cacheprovider

The stand-in answers the C-ABI entry points with the oracle (rhs!, sources, BC passes, SSPRK steps, limiter) and with the
host emulation of the setup kernels (tests/emu).  It says NOTHING about the CUDA code: a test that passes here has only shown
that its script and its reference values are sound.  Not every entry point is covered (ELL operator input, norms field and
the error-path tests of tests/test_gpu_parity.py are not); multi-rank is not supported."""
import ctypes as C
import math
import sys

import numpy as np
import scipy.sparse as sp

import os

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [_ROOT, os.path.join(_ROOT, 'oracle'), os.path.join(_ROOT, 'tests')]
import mft_oracle as orc
import emu


def _addr(x):
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    return C.cast(x, C.c_void_p).value or 0


def _arr(x, n, dtype=np.float64):
    a = _addr(x)
    assert a, "NULL pointer"
    ct = {np.float64: C.c_double, np.int64: C.c_int64, np.int32: C.c_int32}[dtype]
    return np.ctypeslib.as_array(C.cast(a, C.POINTER(ct)), shape=(int(n),))


def _soa(x, V, n):
    ptrs = C.cast(_addr(x), C.POINTER(C.c_void_p))
    return [_arr(ptrs[v], n) for v in range(V)]


class Ctx:
    pass


class MockLib:
    def __init__(self):
        self.ctxs = {}
        self.err = b""
        self.next = 1
        self.real = C.CDLL(os.path.join(_ROOT, 'meshfreetrixi.jl_b200', 'libmft_b200.so'))   # host-only helpers (mft_sfc_order)

    # --- misc
    def mft_last_error(self):
        return self.err

    def mft_device_count(self):
        return 1

    def mft_version(self):
        return 110

    def fail(self, code, msg):
        self.err = msg.encode()
        return code

    def mft_sfc_order(self, n, x, y, out):
        f = self.real.mft_sfc_order
        f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        return f(n, x, y, out)

    def mft_launch_count(self, ctx):
        return 42

    def mft_synchronize(self, ctx):
        c = self._c(ctx)
        if getattr(c, "norm_misses", 0) > getattr(c, "norm_misses_seen", 0):   # loud once, like check_norm_misses
            c.norm_misses_seen = c.norm_misses
            return self.fail(-6, "fused step: rows exceeded the one-pass ode_maximum statistic (mock)")
        return 0

    # --- ctx
    def mft_ctx_create(self, out, device, n_local, n_halo, nvars, ndims, k):
        c = Ctx()
        c.n, c.nh, c.V, c.k = int(n_local), int(n_halo), int(nvars), int(k)
        assert c.nh == 0, "mock: single rank only"
        c.eq = None
        c.ops = {}
        c.bcs, c.srcs = [], []
        c.P = None
        c.u = np.zeros((c.V, c.n))
        c.kf = None
        c.have_fsal = False
        c.nbr = None
        c.stage_lim = None
        c.pending = None
        c.opts = {}
        c.time_map = {}
        h = self.next
        self.next += 1
        self.ctxs[h] = c
        out._obj.value = h
        return 0

    def _c(self, ctx):
        return self.ctxs[ctx.value if hasattr(ctx, "value") else ctx]

    def mft_ctx_destroy(self, ctx):
        self.ctxs.pop(ctx.value if hasattr(ctx, "value") else ctx, None)
        return 0

    def mft_set_equation(self, ctx, kind, prm, nprm):
        c = self._c(ctx)
        c.eq = (int(kind), list(_arr(prm, nprm)))
        return 0

    def mft_set_option(self, ctx, opt, val):
        self._c(ctx).opts[int(opt)] = float(val)
        return 0

    def mft_set_permutation(self, ctx, perm):
        return 0

    def mft_set_order_keys(self, ctx, keys):
        return 0

    def mft_set_operator_csc(self, ctx, slot, cp, rv, nz):
        c = self._c(ctx)
        colptr = _arr(cp, c.n + 1, np.int64).copy()
        nnz = int(colptr[-1] - 1)
        rowval = _arr(rv, nnz, np.int64).copy()
        nzval = _arr(nz, nnz).copy()
        c.ops[int(slot)] = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(c.n, c.n))
        return 0

    def mft_add_boundary(self, ctx, kind, nb, idx1, nrm, vals):
        c = self._c(ctx)
        nb = int(nb)
        idx = _arr(idx1, nb, np.int64).copy() - 1 if nb else np.zeros(0, np.int64)
        normals = _arr(nrm, 2 * nb).reshape(nb, 2).copy() if nb else np.zeros((0, 2))
        values = None
        if int(kind) == 0:
            values = _arr(vals, c.V * nb).reshape(c.V, nb).copy() if nb else np.zeros((c.V, 0))
        okind = {0: orc.BC_DIRICHLET, 1: orc.BC_SLIP_WALL, 2: orc.BC_DO_NOTHING}[int(kind)]
        bc = orc.OracleBC(okind, idx, normals, values=values)
        bc.slots = {}
        if okind == orc.BC_DIRICHLET:
            def fn(x, tt, bc=bc, c=c):
                tab = c.time_map.get(tt)
                if tab is not None and tab in bc.slots:
                    bc.values = bc.slots[tab]      # the step made this slot's table the current one
                return bc.values
            bc.value_fn = fn
        c.bcs.append(bc)
        return 0

    def mft_set_stage_boundary_values(self, ctx, g, slot, vals):
        c = self._c(ctx)
        bc = c.bcs[int(g)]
        if int(slot) not in (0, 1):
            return self.fail(-1, "slot")
        bc.slots[int(slot)] = _arr(vals, c.V * len(bc.idx)).reshape(c.V, len(bc.idx)).copy()
        return 0

    def mft_update_boundary_values(self, ctx, g, vals):
        c = self._c(ctx)
        bc = c.bcs[int(g)]
        bc.values = _arr(vals, c.V * len(bc.idx)).reshape(c.V, len(bc.idx)).copy()
        return 0

    def mft_add_source(self, ctx, kind, prm, nprm, cp, rv, nz):
        c = self._c(ctx)
        kind = int(kind)
        p = list(_arr(prm, nprm))
        if kind in (0, 1):
            colptr = _arr(cp, c.n + 1, np.int64).copy()
            nnz = int(colptr[-1] - 1)
            H = sp.csc_matrix((_arr(nz, nnz).copy(), _arr(rv, nnz, np.int64).copy() - 1, colptr - 1), shape=(c.n, c.n))
            s = orc.OracleSource(kind=orc.SRC_HV_FLYER if kind == 0 else orc.SRC_HV_TOMINEC, hv=orc.JuliaCSC(H), gamma=p[0])
        elif kind == 2:
            s = orc.source_upwind(p[1], c_uw=p[0])
        elif kind == 3:
            s = orc.source_residual(p[2], c_rv=p[0], c_uw=p[1], polydeg=int(p[3]),
                                    mean_divisor_vn=bool(c.opts.get(1, 1.0)), max_lexicographic=bool(c.opts.get(2, 1.0)))
        elif kind == 4:
            if c.eq[0] != 0:
                return self.fail(-3, "mft_add_source: the IGR source is defined for Euler 2-D only")
            s = orc.source_igr(alpha=p[0], maxiter=int(p[1]) if len(p) > 1 else 20)
        else:
            return self.fail(-1, "unknown kind")
        c.srcs.append(s)
        return 0

    def mft_finalize(self, ctx):
        c = self._c(ctx)
        if c.P is None:
            eqk = orc.EQ_EULER2D if c.eq[0] == 0 else orc.EQ_ADVECTION2D
            c.P = orc.OracleProblem(np.zeros((c.n, 2)), c.V, eqk, c.eq[1], c.ops[0], c.ops[1], c.bcs, c.srcs)
        return 0

    # --- compute
    def _get(self, soa, c):
        return np.ascontiguousarray(np.stack(_soa(soa, c.V, c.n)))

    def _put(self, soa, c, a):
        for v, dst in enumerate(_soa(soa, c.V, c.n)):
            dst[:] = a[v]

    def mft_rhs(self, ctx, t, u_soa, du_soa, mem):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        u = self._get(u_soa, c)
        du = c.P.rhs(u, float(t))
        c.u = u.copy()
        c.have_fsal = False
        self._put(u_soa, c, u)
        self._put(du_soa, c, du)
        return 0

    def mft_calc_fluxes(self, ctx, u_soa, du_soa):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        u, du = self._get(u_soa, c), self._get(du_soa, c)
        c.P.calc_fluxes(u, du)
        self._put(du_soa, c, du)
        return 0

    def mft_apply_source(self, ctx, i, t, u_soa, du_soa):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        u, du = self._get(u_soa, c), self._get(du_soa, c)
        c.P.apply_source(int(i), u, du, float(t))
        self._put(du_soa, c, du)
        return 0

    def mft_boundary_pass(self, ctx, t, u_soa, du_soa):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        u, du = self._get(u_soa, c), self._get(du_soa, c)
        c.P.boundary_pass(u, du, float(t))
        self._put(u_soa, c, u)
        self._put(du_soa, c, du)
        return 0

    def mft_upload_state(self, ctx, soa):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        c.u = self._get(soa, c)
        c.have_fsal = False
        c.nsteps_since_upload = 0
        return 0

    def mft_download_state(self, ctx, soa):
        c = self._c(ctx)
        self._put(soa, c, c.u)
        return 0

    def mft_history_push(self, ctx, t, success_iter, approx_order):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        c.P.history_callback(c.u, float(t), int(success_iter), int(approx_order))
        return 0

    def _limit(self, c, u):
        if c.stage_lim:
            thr, var = c.stage_lim
            orc.limiter_zhang_shu(u, c.nbr, thr, var, c.eq[1][0])

    def mft_ssprk_step(self, ctx, scheme, t, dt):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        # MFT_MOCK_NORM_MISSES="step,count": pretend that the fused step's one-pass norms missed `count` rows in the step-th
        # mft_ssprk_step call since the last mft_upload_state (the two-pass kernels, MFT_OPT_FUSED_STEP = 0, never miss): lets a
        # CPU test drive the miss handling of bench.py
        c.nsteps_since_upload = getattr(c, "nsteps_since_upload", 0) + 1
        spec = os.environ.get("MFT_MOCK_NORM_MISSES")
        if spec and c.opts.get(12, 1.0) != 0.0:
            step, count = (int(x) for x in spec.split(","))
            if c.nsteps_since_upload == step:
                c.norm_misses = getattr(c, "norm_misses", 0) + count
        L = orc.lib()
        t, dt = float(t), float(dt)
        u = c.u
        if not c.have_fsal:
            c.kf = c.P.rhs(u, t)
            c.have_fsal = True
        uprev = u.copy()
        c.time_map = {t + dt: 0, t + dt / 2: 1}
        for s, ts in ((1, t + dt), (2, t + dt / 2), (3, t + dt)):
            L.orc_ssprk33_stage(C.c_int64(u.size), s, C.c_double(dt), C.c_void_p(uprev.ctypes.data), C.c_void_p(c.kf.ctypes.data),
                                C.c_void_p(u.ctypes.data))
            self._limit(c, u)
            c.kf = c.P.rhs(u, ts)
        c.time_map = {}
        return 0

    def mft_ssprk43_step(self, ctx, t, dt, abstol, reltol, ss_out, cnt_out):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        if c.stage_lim:
            return self.fail(-3, "mft_ssprk43_step: the stage limiter is wired into mft_ssprk_step (SSPRK33) only")
        if not c.have_fsal:
            c.kf = c.P.rhs(c.u, float(t))
            c.have_fsal = True
        c.time_map = {float(t) + float(dt): 0, float(t) + float(dt) / 2: 1}
        un, kn, eest = orc._ssprk43_step(c.P, c.u, c.kf, float(t), float(dt), float(abstol), float(reltol))
        c.time_map = {}
        c.pending = (c.u.copy(), c.kf.copy())
        ss_out._obj.value = eest * eest * c.u.size
        cnt_out._obj.value = c.u.size
        c.u, c.kf = un, kn
        return 0

    def mft_step_commit(self, ctx, accept):
        c = self._c(ctx)
        if not int(accept):
            c.u, c.kf = c.pending
        c.pending = None
        return 0

    def mft_get_field(self, ctx, field, out):
        c = self._c(ctx)
        names = {0: "eps", 1: "eps_uw", 2: "eps_rv", 3: "eps_c", 4: "residual", 5: "approx_du"}
        field = int(field)
        if field in names:
            for s in c.srcs:
                if names[field] in s.arrays:
                    if not c.opts.get(3, 0.0) and field != 5:
                        return self.fail(-1, "mft_get_field: field not available (enable MFT_OPT_DIAGNOSTICS)")
                    a = np.asarray(s.arrays[names[field]], dtype=np.float64).reshape(-1)
                    _arr(out, a.size)[:] = a
                    return 0
            return self.fail(-1, "no such field")
        if field == 7:
            for s in c.srcs:
                if s.kind == orc.SRC_IGR:
                    _arr(out, c.n)[:] = s.arrays["sigma"]
                    return 0
            return self.fail(-1, "mft_get_field: no IGR source")
        if field == 8:
            for s in c.srcs:
                if s.kind == orc.SRC_IGR:
                    o = _arr(out, 3)
                    o[0], o[1], o[2] = s.arrays["iters"], s.arrays["res"], s.arrays.get("res0", s.arrays["res"])
                    return 0
            return self.fail(-1, "mft_get_field: no IGR source")
        if field == 6:
            _arr(out, c.V)[:] = 0.0
            return 0
        if field == 9:   # MFT_FIELD_NORM_MISSES: the stand-in has no one-pass statistic (see MFT_MOCK_NORM_MISSES)
            _arr(out, 1)[:] = float(getattr(c, "norm_misses", 0))
            return 0
        return self.fail(-1, "bad field")

    def mft_count_nonfinite(self, ctx, out):
        c = self._c(ctx)
        out._obj.value = int((~np.isfinite(c.u)).sum())
        return 0

    # --- limiter
    def mft_set_neighbors(self, ctx, nbr1):
        c = self._c(ctx)
        c.nbr = _arr(nbr1, c.n * c.k, np.int64).reshape(c.n, c.k).copy() - 1
        return 0

    def mft_limiter_zhang_shu(self, ctx, npairs, thr, var, u_soa, mem):
        c = self._c(ctx)
        self.mft_finalize(ctx)
        u = self._get(u_soa, c)
        orc.limiter_zhang_shu(u, c.nbr, list(_arr(thr, npairs)), [int(v) for v in _arr(var, npairs, np.int32)], c.eq[1][0])
        self._put(u_soa, c, u)
        c.u = u.copy()
        c.have_fsal = False
        return 0

    def mft_set_stage_limiter(self, ctx, npairs, thr, var):
        c = self._c(ctx)
        npairs = int(npairs)
        c.stage_lim = (list(_arr(thr, npairs)), [int(v) for v in _arr(var, npairs, np.int32)]) if npairs else None
        return 0

    # --- setup pipeline: the emulated kernels
    def mft_setup_knn(self, dev, n, x, y, k, nbr, dist):
        pts = np.stack([_arr(x, n), _arr(y, n)], axis=1)
        try:
            nb, d = emu.setup_knn(pts, int(k))
        except emu.EmuError as e:
            return self.fail(-1, str(e))
        _arr(nbr, n * k, np.int64)[:] = (nb + 1).reshape(-1)
        if _addr(dist):
            _arr(dist, n * k)[:] = d.reshape(-1)
        return 0

    def mft_setup_knn_queries(self, dev, n, x, y, k, nq, q1, nbr, dist):
        pts = np.stack([_arr(x, n), _arr(y, n)], axis=1)
        q = _arr(q1, nq, np.int64) - 1
        try:
            nb, d = emu.setup_knn(pts, int(k), queries=q)
        except emu.EmuError as e:
            return self.fail(-1, str(e))
        _arr(nbr, nq * k, np.int64)[:] = (nb + 1).reshape(-1)
        _arr(dist, nq * k)[:] = d.reshape(-1)
        return 0

    def _weights(self, n, x, y, n_rows, k, nbr1, p, N, kk, wx, wy, hyb=None):
        pts = np.stack([_arr(x, n), _arr(y, n)], axis=1)
        nb = _arr(nbr1, n_rows * k, np.int64).reshape(n_rows, k) - 1
        try:
            ex, ey = emu.setup_rbf_weights(pts, nb, int(p), int(N), int(kk), hybrid=hyb)
        except emu.EmuError as e:
            return self.fail(-1, str(e))
        _arr(wx, n_rows * k)[:] = ex.reshape(-1)
        _arr(wy, n_rows * k)[:] = ey.reshape(-1)
        return 0

    def mft_setup_rbf_weights(self, dev, n, x, y, k, nbr1, p, N, kk, wx, wy):
        return self._weights(n, x, y, n, k, nbr1, p, N, kk, wx, wy)

    def mft_setup_rbf_weights_rows(self, dev, n, x, y, n_rows, k, nbr1, p, N, kk, wx, wy):
        return self._weights(n, x, y, n_rows, k, nbr1, p, N, kk, wx, wy)

    def mft_setup_rbf_weights_hybrid(self, dev, n, x, y, n_rows, k, nbr1, p, a, b, e, N, kk, wx, wy):
        return self._weights(n, x, y, n_rows, k, nbr1, p, N, kk, wx, wy, hyb=(a, b, e))


def install():
    import mft_b200 as m

    mock = MockLib()
    m._lib.load = lambda: mock
    m._lib._lib = mock
    return mock


import time as _time
_keep = []


def _timer_start(self, ctx):
    self._t0 = _time.perf_counter()
    return 0


def _timer_stop(self, ctx, out):
    out._obj.value = (_time.perf_counter() - self._t0) * 1e3
    return 0


def _set_kernel_timing(self, ctx, on):
    return 0


def _kernel_time_ms(self, ctx, cls, ms, cnt):
    ms._obj.value = 1.0 + cls
    cnt._obj.value = 3
    return 0


def _host_alloc(self, out, n):
    buf = (C.c_char * int(n))()
    _keep.append(buf)
    out._obj.value = C.addressof(buf)
    return 0


MockLib.mft_timer_start = _timer_start
MockLib.mft_timer_stop = _timer_stop
MockLib.mft_set_kernel_timing = _set_kernel_timing
MockLib.mft_kernel_time_ms = _kernel_time_ms
MockLib.mft_host_alloc = _host_alloc
