"""Synthetic iNMF / UINMF inputs of the shapes BASELINE.json names.

Planted low-rank Poisson counts: A ~ Gamma(0.5) (m x r, shared by all datasets), per cell block
B ~ Gamma(0.5) (r x cells), C ~ Poisson(s * A @ B) with s bisected so that the expected density is
`density` (default 5 %).  Then, exactly as rliger's `normalize` (reference R/preprocess.R:592-608) and
`scaleNotCenter` (src/data_processing.cpp:42-58): divide every column by its library size and every
row by sqrt(sum(x^2) / (n - 1)); explicit zeros are dropped; output is CSC (genes x cells).

Two generators with the same recipe:
  * make_numpy : numpy default_rng([20261019, i, b]) per block -- small test cases, runs anywhere;
  * make_torch : torch on the current CUDA device -- the bench-size configurations (1M cells in seconds).
This is input generation only; no part of the factorisation lives here.
"""
import numpy as np
import scipy.sparse as sp

SEED = 20261019


def _bisect_scale(mean_fn, density):
    lo, hi = 1e-8, 10.0
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if mean_fn(mid) < density:
            lo = mid
        else:
            hi = mid
    return float(np.sqrt(lo * hi))


def make_numpy(d, n_per, m, r, density=0.05, u=0, block=5000, seed=SEED, return_scales=False, return_unshared_scales=False):
    """Returns (Es, Ps): lists of scipy CSC float64 (m x n_i) and (u x n_i) or None.
    return_scales: also returns [(row_scale, col_scale)] per dataset for the shared block: E = counts * row_scale[:, None] *
    col_scale[None, :] (1 / row rms, 1 / library size) -- the count structure liger_b200.CountScaled carries.
    return_unshared_scales: a fourth list with the same pair for the unshared blocks (own row rms, the SAME library sizes)."""
    A = np.random.default_rng([seed, 999]).gamma(0.5, 1.0, (m, r))
    Au = [np.random.default_rng([seed, 999, i]).gamma(0.5, 1.0, (u, r)) if u else None for i in range(d)]
    B0 = np.random.default_rng([seed, 0, 0]).gamma(0.5, 1.0, (r, min(block, n_per)))
    rate0 = A @ B0
    s = _bisect_scale(lambda x: float(np.mean(1.0 - np.exp(-x * rate0))), density)
    Es, Ps, scales, pscales = [], [], [], []
    for i in range(d):
        cols_e, cols_p = [], []
        for b, c0 in enumerate(range(0, n_per, block)):
            nb = min(block, n_per - c0)
            rng = np.random.default_rng([seed, i, b])
            B = rng.gamma(0.5, 1.0, (r, nb))
            C = rng.poisson(s * (A @ B)).astype(np.float64)
            if u:
                Cu = rng.poisson(s * (Au[i] @ B)).astype(np.float64)
            # guard against empty cells (library size 0): give them one count in gene 0
            empty = C.sum(axis=0) == 0
            C[0, empty] = 1.0
            cols_e.append(sp.csc_matrix(C))
            if u:
                cols_p.append(sp.csc_matrix(Cu))
        E = sp.hstack(cols_e, format="csc")
        lib = np.asarray(E.sum(axis=0)).ravel()
        Xs, rms = _scale(E, lib, True)
        Es.append(Xs)
        scales.append((1.0 / rms, 1.0 / lib))
        if u:
            P = sp.hstack(cols_p, format="csc")
            Pscaled, prms = _scale(P, lib, True)
            Ps.append(Pscaled)
            pscales.append((1.0 / prms, 1.0 / lib))
        else:
            Ps.append(None)
            pscales.append(None)
    if return_unshared_scales:
        return Es, Ps, scales, pscales
    if return_scales:
        return Es, Ps, scales
    return Es, Ps


def _scale(X, lib, return_rms=False):
    X = sp.csc_matrix(X, dtype=np.float64)
    X.data = X.data / np.repeat(lib, np.diff(X.indptr))
    n = X.shape[1]
    Xr = sp.csr_matrix(X)
    ss = np.asarray(Xr.multiply(Xr).sum(axis=1)).ravel() / (n - 1)
    rms = np.sqrt(ss)
    rms[rms == 0] = 1.0
    Xr.data = Xr.data / np.repeat(rms, np.diff(Xr.indptr))
    out = sp.csc_matrix(Xr)
    out.eliminate_zeros()
    out.sort_indices()
    if return_rms:
        return out, rms
    return out


def make_torch(d, n_per, m, r, density=0.05, u=0, block=25000, seed=SEED, device=None, pin=True, value_dtype="float64",
               shard=0):
    """GPU generator.  Returns (Es, Ps, info) where each entry is a tuple
    (indptr int32, indices int32, data, shape) of (pinned) host numpy arrays -- the dgCMatrix slots."""
    import torch
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    g = torch.Generator(device=dev)
    g.manual_seed(seed)

    def gamma(shape):
        return torch._standard_gamma(torch.full(shape, 0.5, device=dev, dtype=torch.float32), generator=g)

    A = gamma((m, r))                      # gene loadings: identical on every shard
    Au = [gamma((u, r)) if u else None for _ in range(d)]
    rate0 = A @ gamma((r, min(block, n_per)))
    s = _bisect_scale(lambda x: float(torch.mean(1.0 - torch.exp(-x * rate0)).item()), density)
    g.manual_seed(seed + 7919 * (shard + 1))   # cells: different on every shard (rank)
    vdt = np.float64 if value_dtype == "float64" else np.float32

    def host(t, dtype):
        t = t.to(dtype)
        if pin and dev.type == "cuda":
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        return t.cpu()

    def one(Amat, nrows, i, Bs_out, lib_in):
        # returns (indptr, rows, vals, lib): CSC pieces on the device, before row scaling
        counts_per_col, rows_l, vals_l = [], [], []
        sumsq = torch.zeros(nrows, device=dev, dtype=torch.float64)
        libs = []
        for b, c0 in enumerate(range(0, n_per, block)):
            nb = min(block, n_per - c0)
            if Bs_out is not None and len(Bs_out) > b:
                B = Bs_out[b]
            else:
                B = gamma((r, nb))
                if Bs_out is not None:
                    Bs_out.append(B)
            C = torch.poisson(s * (Amat @ B), generator=g)           # nrows x nb
            if lib_in is None:
                lib = C.sum(dim=0)
                empty = lib == 0
                if bool(empty.any()):
                    C[0, empty] = 1.0
                    lib = C.sum(dim=0)
            else:
                lib = lib_in[b]
            libs.append(lib)
            Ct = C.t().contiguous()                                     # nb x nrows  (column-major walk)
            nzc, nzr = Ct.nonzero(as_tuple=True)
            v = (Ct[nzc, nzr] / lib[nzc]).to(torch.float64)
            sumsq.index_add_(0, nzr, v * v)
            counts_per_col.append(torch.bincount(nzc, minlength=nb))
            rows_l.append(nzr.to(torch.int32))
            vals_l.append(v)
            del C, Ct
        rows = torch.cat(rows_l)
        vals = torch.cat(vals_l)
        cnt = torch.cat(counts_per_col)
        rms = torch.sqrt(sumsq / (n_per - 1))
        rms[rms == 0] = 1.0
        vals = vals / rms[rows.long()]
        indptr = torch.zeros(n_per + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(cnt, 0)
        return indptr, rows, vals, libs, rms

    Es, Ps, scales, pscales = [], [], [], []
    nnz = 0
    for i in range(d):
        Bs = []
        indptr, rows, vals, libs, rms = one(A, m, i, Bs, None)
        nnz += int(rows.numel())
        scales.append(((1.0 / rms).cpu().numpy(), (1.0 / torch.cat(libs).to(torch.float64)).cpu().numpy()))
        Es.append((host(indptr, torch.int32).numpy(), host(rows, torch.int32).numpy(),
                   host(vals, torch.float64 if vdt is np.float64 else torch.float32).numpy(), (m, n_per)))
        if u:
            indptr, rows, vals, _, prms = one(Au[i], u, i, Bs, libs)
            nnz += int(rows.numel())
            pscales.append(((1.0 / prms).cpu().numpy(), scales[-1][1]))     # own row rms, the shared block's library sizes
            Ps.append((host(indptr, torch.int32).numpy(), host(rows, torch.int32).numpy(),
                       host(vals, torch.float64 if vdt is np.float64 else torch.float32).numpy(), (u, n_per)))
        else:
            Ps.append(None)
            pscales.append(None)
        del Bs
    if dev.type == "cuda":
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return Es, Ps, {"scale": s, "nnz": nnz, "density": nnz / float(d * n_per * (m + u)), "scales": scales,
                    "unshared_scales": pscales}


def to_scipy(t):
    indptr, indices, data, shape = t
    return sp.csc_matrix((np.asarray(data, dtype=np.float64), indices, indptr), shape=shape)
