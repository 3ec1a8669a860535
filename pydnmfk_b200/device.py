"""Device-side operator layer: torch CUDA tensors in, libdnmf kernels underneath.

PyTorch is used only to own device memory and streams (and, in dist_comm, for
the process groups).  Every numeric operation of the update loop is one of the
hand-written kernels behind the C-ABI (include/dnmf.h); there is no CPU or
eager-PyTorch fallback -- calling this layer without a CUDA device raises.
"""
import numpy as np
import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.float64: L.F64}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('pydnmfk_b200 needs a CUDA device (sm_100a); it has no CPU fallback')


def to_device(x, dtype=None, device=None):
    """numpy / torch (host or device) -> contiguous CUDA tensor."""
    require_cuda()
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    t = t.to(device, non_blocking=False)
    return t.contiguous()


def to_device_view(x, dtype=None):
    """Like to_device but keeps the strides of a (possibly transposed) device tensor / numpy view."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda()
    else:
        a = np.asarray(x)
        t = torch.from_numpy(a.copy() if not a.flags.writeable else a).cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def torch_dtype(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float32:
        return torch.float32
    if np_dtype == np.float64:
        return torch.float64
    raise TypeError('precision %s is not supported on the device path (float32/float64 only)' % np_dtype)


def _ld(t):
    """Leading dimension of a row-major 2-D tensor (elements)."""
    assert t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1), 'row-major matrix expected'
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


class DeviceOps:
    """Kernel launcher bound to one device; owns the scratch workspace."""

    def __init__(self, device=None, math_mode=L.MATH_ACCURATE):
        require_cuda()
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.math_mode = math_mode
        self._ws = None
        self._ws_bytes = 0
        self.timers = None      # set to {} to bracket the A-streaming passes with CUDA events (bench.py)

    def _t0(self, name):
        if self.timers is None:
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(torch.cuda.current_stream(self.device))
        self.timers.setdefault(name, []).append(ev)
        return ev

    def _t1(self, ev):
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))

    def timer_summary(self):
        """{op: (launches, mean ms)} from the recorded events (call after a synchronize)."""
        out = {}
        for name, evs in (self.timers or {}).items():
            ms = [a.elapsed_time(b) for a, b in evs]
            out[name] = (len(ms), sum(ms) / max(len(ms), 1))
        return out

    # ---- plumbing -------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def workspace(self, nbytes):
        if nbytes > self._ws_bytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws_bytes = int(nbytes)
        return self._ws.data_ptr() if self._ws is not None else 0

    def _ws_for(self, op, m, n, k, dt):
        nb = L.workspace_bytes(op, m, n, k, dt)
        return self.workspace(nb), max(nb, self._ws_bytes)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ---- A-streaming contractions -----------------------------------------------------------
    def ah(self, A, H, out=None):
        """V = A @ H.T   (dist_nmf.py:198, :730)"""
        m, n = A.shape
        k = H.shape[0]
        dt = _DT[A.dtype]
        V = out if out is not None else self.empty((m, k), A.dtype)
        ws, wsb = self._ws_for(L.OP_AH, m, n, k, dt)
        ev = self._t0('ah')
        L.call('dnmf_ah', A.data_ptr(), _ld(A), H.data_ptr(), _ld(H), V.data_ptr(), _ld(V), m, n, k, dt,
               self.math_mode, ws, wsb, self._stream())
        self._t1(ev)
        return V

    def wta(self, A, W, transposed_out=False, out=None):
        """Y = W.T @ A (k x n), or Y.T (n x k) when transposed_out   (dist_nmf.py:166, :749)"""
        m, n = A.shape
        k = W.shape[1]
        dt = _DT[A.dtype]
        Y = out if out is not None else self.empty((n, k) if transposed_out else (k, n), A.dtype)
        ws, wsb = self._ws_for(L.OP_WTA, m, n, k, dt)
        ev = self._t0('wta')
        L.call('dnmf_wta', A.data_ptr(), _ld(A), W.data_ptr(), _ld(W), Y.data_ptr(), _ld(Y), m, n, k,
               1 if transposed_out else 0, dt, self.math_mode, ws, wsb, self._stream())
        self._t1(ev)
        return Y

    def kl_uht(self, A, W, H, eps, out=None):
        """V = (A / (W @ H + eps)) @ H.T without materialising W @ H   (dist_nmf.py:338-339, :806,:810)"""
        m, n = A.shape
        k = H.shape[0]
        dt = _DT[A.dtype]
        V = out if out is not None else self.empty((m, k), A.dtype)
        ws, wsb = self._ws_for(L.OP_KL_UHT, m, n, k, dt)
        ev = self._t0('kl_uht')
        L.call('dnmf_kl_uht', A.data_ptr(), _ld(A), W.data_ptr(), _ld(W), H.data_ptr(), _ld(H), V.data_ptr(),
               _ld(V), m, n, k, float(eps), dt, self.math_mode, ws, wsb, self._stream())
        self._t1(ev)
        return V

    def kl_wtu(self, A, W, H, eps, transposed_out=False, out=None):
        """Y = W.T @ (A / (W @ H + eps))   (dist_nmf.py:312-313, :806,:808)"""
        m, n = A.shape
        k = H.shape[0]
        dt = _DT[A.dtype]
        Y = out if out is not None else self.empty((n, k) if transposed_out else (k, n), A.dtype)
        ws, wsb = self._ws_for(L.OP_KL_WTU, m, n, k, dt)
        ev = self._t0('kl_wtu')
        L.call('dnmf_kl_wtu', A.data_ptr(), _ld(A), W.data_ptr(), _ld(W), H.data_ptr(), _ld(H), Y.data_ptr(),
               _ld(Y), m, n, k, float(eps), 1 if transposed_out else 0, dt, self.math_mode, ws, wsb,
               self._stream())
        self._t1(ev)
        return Y

    def gram(self, X, trans):
        """trans=False: X.T @ X for X [rows x k];  trans=True: X @ X.T for X [k x rows]."""
        if trans:
            k, rows = X.shape
        else:
            rows, k = X.shape
        dt = _DT[X.dtype]
        G = self.empty((k, k), X.dtype)
        ws, wsb = self._ws_for(L.OP_GRAM, rows, 0, k, dt)
        L.call('dnmf_gram', X.data_ptr(), _ld(X), rows, k, 1 if trans else 0, G.data_ptr(), dt, ws, wsb,
               self._stream())
        return G

    # ---- updates ----------------------------------------------------------------------------
    @staticmethod
    def _ystrides(Y, transposed):
        # element (kk, c) of the k x n quantity; Y may be stored as its transpose (n x k)
        return (1, _ld(Y)) if transposed else (_ld(Y), 1)

    def mu_update_w(self, W, V, G, eps):
        m, k = W.shape
        L.call('dnmf_mu_update_w', W.data_ptr(), _ld(W), V.data_ptr(), _ld(V), G.data_ptr(), m, k, float(eps),
               _DT[W.dtype], self._stream())

    def mu_update_h(self, H, Y, G, eps, y_transposed=False, clamp=False):
        k, n = H.shape
        sk, sc = self._ystrides(Y, y_transposed)
        L.call('dnmf_mu_update_h', H.data_ptr(), _ld(H), Y.data_ptr(), sk, sc, G.data_ptr(), k, n, float(eps),
               1 if clamp else 0, _DT[H.dtype], self._stream())

    def kl_update_w(self, W, V, x, eps):
        m, k = W.shape
        L.call('dnmf_kl_update_w', W.data_ptr(), _ld(W), V.data_ptr(), _ld(V), x.data_ptr(), m, k, float(eps),
               _DT[W.dtype], self._stream())

    def kl_update_h(self, H, Y, x, eps, y_transposed=False, clamp=False):
        k, n = H.shape
        sk, sc = self._ystrides(Y, y_transposed)
        L.call('dnmf_kl_update_h', H.data_ptr(), _ld(H), Y.data_ptr(), sk, sc, x.data_ptr(), k, n, float(eps),
               1 if clamp else 0, _DT[H.dtype], self._stream())

    def clamp_min(self, X, lo):
        r, c = X.shape
        L.call('dnmf_clamp_min', X.data_ptr(), _ld(X), r, c, float(lo), _DT[X.dtype], self._stream())

    # ---- small reductions ---------------------------------------------------------------------
    def colsum(self, X):
        r, c = X.shape
        out = self.empty((c,), X.dtype)
        ws, wsb = self._ws_for(L.OP_SUMS, r, c, c, _DT[X.dtype])
        L.call('dnmf_colsum', X.data_ptr(), _ld(X), r, c, out.data_ptr(), _DT[X.dtype], ws, wsb, self._stream())
        return out

    def rowsum(self, X):
        r, c = X.shape
        out = self.empty((r,), X.dtype)
        ws, wsb = self._ws_for(L.OP_SUMS, r, c, r, _DT[X.dtype])
        L.call('dnmf_rowsum', X.data_ptr(), _ld(X), r, c, out.data_ptr(), _DT[X.dtype], ws, wsb, self._stream())
        return out

    def sqnorm(self, X):
        """sum(X**2) as a float64 device scalar (shape [1])."""
        r, c = X.shape
        out = self.empty((1,), torch.float64)
        ws, wsb = self._ws_for(L.OP_SUMS, r, c, min(c, L.MAX_K), _DT[X.dtype])
        L.call('dnmf_sqnorm', X.data_ptr(), _ld(X), r, c, out.data_ptr(), _DT[X.dtype], ws, wsb, self._stream())
        return out

    def normalize(self, W, H, s, eps):
        m, k = W.shape
        n = H.shape[1]
        L.call('dnmf_normalize', W.data_ptr(), _ld(W), m, H.data_ptr(), _ld(H), n, k, s.data_ptr(), float(eps),
               _DT[W.dtype], self._stream())

    def residual_sqnorm(self, A, W, H):
        """[||A - W H||_F^2, ||A||_F^2] as a float64 device tensor of shape [2]."""
        m, n = A.shape
        k = W.shape[1]
        dt = _DT[A.dtype]
        out = self.empty((2,), torch.float64)
        ws, wsb = self._ws_for(L.OP_RESIDUAL, m, n, k, dt)
        L.call('dnmf_residual_sqnorm', A.data_ptr(), _ld(A), W.data_ptr(), _ld(W), H.data_ptr(), _ld(H), m, n, k,
               out.data_ptr(), dt, ws, wsb, self._stream())
        return out

    def column_err(self, A, W, H):
        m, n = A.shape
        k = W.shape[1]
        num = self.empty((n,), torch.float64)
        den = self.empty((n,), torch.float64)
        L.call('dnmf_column_err', A.data_ptr(), _ld(A), W.data_ptr(), _ld(W), H.data_ptr(), _ld(H), m, n, k,
               num.data_ptr(), den.data_ptr(), _DT[A.dtype], self._stream())
        return num, den

    # ---- HALS / BCD ---------------------------------------------------------------------------
    def hals_w_col(self, W, V, G, kk, eps):
        m, k = W.shape
        sq = self.empty((1,), torch.float64)
        ws, wsb = self._ws_for(L.OP_SUMS, m, k, k, _DT[W.dtype])
        L.call('dnmf_hals_w_col', W.data_ptr(), _ld(W), V.data_ptr(), _ld(V), G.data_ptr(), m, k, kk, float(eps),
               sq.data_ptr(), _DT[W.dtype], ws, wsb, self._stream())
        return sq

    def div_col(self, W, kk, ss_sq):
        L.call('dnmf_div_col', W.data_ptr(), _ld(W), W.shape[0], kk, ss_sq.data_ptr(), _DT[W.dtype], self._stream())

    def hals_h(self, H, Y, G, eps, y_transposed=False):
        k, n = H.shape
        sk, sc = self._ystrides(Y, y_transposed)
        L.call('dnmf_hals_h', H.data_ptr(), _ld(H), Y.data_ptr(), sk, sc, G.data_ptr(), k, n, float(eps),
               _DT[H.dtype], self._stream())

    def bcd_pg_w(self, W, Wm, V, G, Lip):
        m, k = W.shape
        L.call('dnmf_bcd_pg_w', W.data_ptr(), _ld(W), Wm.data_ptr(), _ld(Wm), V.data_ptr(), _ld(V), G.data_ptr(),
               m, k, float(Lip), _DT[W.dtype], self._stream())

    def bcd_pg_h(self, H, Hm, Y, G, Lip, y_transposed=False):
        k, n = H.shape
        sk, sc = self._ystrides(Y, y_transposed)
        L.call('dnmf_bcd_pg_h', H.data_ptr(), _ld(H), Hm.data_ptr(), _ld(Hm), Y.data_ptr(), sk, sc, G.data_ptr(),
               k, n, float(Lip), _DT[H.dtype], self._stream())

    def div_cols(self, W, s):
        m, k = W.shape
        L.call('dnmf_div_cols', W.data_ptr(), _ld(W), m, k, s.data_ptr(), _DT[W.dtype], self._stream())

    def axpby(self, out, x, y, a, b):
        assert out.is_contiguous() and x.is_contiguous() and y.is_contiguous()
        L.call('dnmf_axpby', out.data_ptr(), x.data_ptr(), y.data_ptr(), float(a), float(b), out.numel(),
               _DT[out.dtype], self._stream())

    # ---- shard ops ------------------------------------------------------------------------------
    def nnz_counts(self, A):
        m, n = A.shape
        rows = self.empty((m,), torch.int64)
        cols = self.empty((n,), torch.int64)
        L.call('dnmf_nnz_counts', A.data_ptr(), _ld(A), m, n, rows.data_ptr(), cols.data_ptr(), _DT[A.dtype],
               self._stream())
        return rows, cols

    def compact(self, A, row_idx, col_idx):
        mr, nc = row_idx.numel(), col_idx.numel()
        out = self.empty((mr, nc), A.dtype)
        L.call('dnmf_compact', A.data_ptr(), _ld(A), row_idx.data_ptr(), mr, col_idx.data_ptr(), nc, out.data_ptr(),
               max(nc, 1), _DT[A.dtype], self._stream())
        return out

    def scatter_rows(self, X, row_idx, total_rows):
        out = torch.zeros((total_rows, X.shape[1]), dtype=torch.float64, device=self.device)
        L.call('dnmf_scatter_rows', X.data_ptr(), _ld(X), row_idx.data_ptr(), X.shape[0], X.shape[1],
               out.data_ptr(), max(X.shape[1], 1), _DT[X.dtype], self._stream())
        return out

    def scatter_cols(self, X, col_idx, total_cols):
        out = torch.zeros((X.shape[0], total_cols), dtype=torch.float64, device=self.device)
        L.call('dnmf_scatter_cols', X.data_ptr(), _ld(X), col_idx.data_ptr(), X.shape[1], X.shape[0],
               out.data_ptr(), max(total_cols, 1), _DT[X.dtype], self._stream())
        return out

    def perturb_uniform(self, A, U, noise_var):
        X = torch.empty_like(A)
        assert A.is_contiguous() and U.is_contiguous()
        L.call('dnmf_perturb_uniform', A.data_ptr(), U.data_ptr(), X.data_ptr(), A.numel(), float(noise_var),
               _DT[A.dtype], self._stream())
        return X


_default_ops = {}


def default_ops():
    require_cuda()
    dev = torch.cuda.current_device()
    if dev not in _default_ops:
        _default_ops[dev] = DeviceOps(torch.device('cuda', dev))
    return _default_ops[dev]
