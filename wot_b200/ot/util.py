"""Local PCA, the producer of the hot path's inputs (reference: wot/ot/util.py:240-255).

`compute_pca` keeps the reference's signature and return tuple.  Where scikit-learn's PCA(svd_solver='auto') would
run its randomized solver (always at atlas shapes: max(shape) > 500 and k < 0.8 min(shape)) the same arithmetic
runs on the GPU (csrc/pca.cu, float64): same centring, same Gaussian test matrix (numpy RandomState(58951)), same
number of power iterations, same sign convention, so the components equal scikit-learn's to roundoff
(tests: cost matrices agree to 1e-9).  Tiny or near-full-rank problems, for which scikit-learn itself switches to an
exact LAPACK solver, and rank-deficient input take the exact GPU path (compute_pca_gpu_exact: Gram matrix of the short
side, symmetric eigendecomposition, float64 on the device).  scikit-learn is never called by the product path.
"""
from __future__ import annotations

import ctypes as C
import logging
import threading

import numpy as np
import scipy.sparse

from .. import _lib

RANDOM_STATE = 58951          # util.py:248
N_OVERSAMPLES = 10            # sklearn PCA default


class LocalPCA:
    """The attributes of the fitted sklearn PCA that the transport-map path reads (ot_model.py:297-301)."""

    def __init__(self, components, singular_values, n_samples, cell_means, gpu_ms):
        self.components_ = components                  # [k, cells]
        self.singular_values_ = singular_values        # [k]
        self.n_components_ = self.n_components = components.shape[0]
        self.explained_variance_ = singular_values ** 2 / (n_samples - 1)
        self.mean_ = cell_means
        self.gpu_ms = gpu_ms


def _sklearn_has_covariance_eigh():
    try:
        import sklearn
        major, minor = (int(v) for v in sklearn.__version__.split(".")[:2])
        return (major, minor) >= (1, 5)
    except Exception:
        return True


def sklearn_solver_choice(n_samples, n_features, k):
    """Which solver the INSTALLED scikit-learn's PCA(svd_solver='auto').fit picks for an (n_samples, n_features)
    matrix (sklearn/decomposition/_pca.py::_fit; the covariance_eigh rule exists from 1.5 on)."""
    if _sklearn_has_covariance_eigh() and n_features <= 1000 and n_samples >= 10 * n_features:
        return "covariance_eigh"
    if max(n_samples, n_features) <= 500:
        return "full"
    if 1 <= k < 0.8 * min(n_samples, n_features):
        return "randomized"
    return "full"


def _dense(m):
    return m.toarray() if scipy.sparse.isspmatrix(m) else np.asarray(m)


def compute_pca_sklearn(m1, m2, n_components):
    """The reference's own arithmetic (util.py:240-255) on scikit-learn."""
    import sklearn.decomposition
    dense = [_dense(m1), _dense(m2)]
    stacked = np.vstack(dense)
    gene_means = stacked.mean(axis=0)
    stacked = stacked - gene_means
    n_components = min(n_components, stacked.shape[0])  # cannot exceed the number of cells
    pca = sklearn.decomposition.PCA(n_components=n_components, random_state=RANDOM_STATE)
    pca.fit(stacked.T)
    loadings = pca.components_.T
    n1 = dense[0].shape[0]
    return loadings[:n1], loadings[n1:n1 + dense[1].shape[0]], pca, gene_means


def compute_pca_gpu(m1, m2, n_components, ctx=None):
    """Randomized-solver path of util.py:240-255 on the GPU (wotb_pca_host)."""
    a = np.ascontiguousarray(_dense(m1), dtype=np.float64)
    b = np.ascontiguousarray(_dense(m2), dtype=np.float64)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("m1 and m2 must be 2-D with the same number of genes")
    n1, n2, genes = a.shape[0], b.shape[0], a.shape[1]
    cells = n1 + n2
    k = min(int(n_components), cells)
    size = k + N_OVERSAMPLES
    n_iter = 7 if k < 0.1 * min(genes, cells) else 4                   # randomized_svd(n_iter='auto')
    short = genes if genes < cells else cells                          # transpose='auto'
    q0 = np.random.RandomState(RANDOM_STATE).normal(size=(short, size))   # randomized_range_finder's test matrix
    comp = np.empty((cells, k))
    sv = np.empty(k)
    gene_means = np.empty(genes)
    cell_means = np.empty(cells)
    ms = C.c_double()
    ctx = ctx or _lib.context()
    _lib.check(ctx.lib.wotb_pca_host(ctx.handle, _lib.ptr(a), n1, _lib.ptr(b), n2, genes, k, _lib.ptr(q0), size, n_iter,
                                     _lib.ptr(comp), _lib.ptr(sv), _lib.ptr(gene_means), _lib.ptr(cell_means),
                                     C.byref(ms)))
    # svd_flip(u_based_decision=False): the largest-magnitude loading of every component is positive
    top = np.argmax(np.abs(comp), axis=0)
    comp *= np.sign(comp[top, np.arange(k)])
    pca = LocalPCA(comp.T, sv, genes, cell_means, ms.value)
    return comp[:n1], comp[n1:], pca, gene_means


EXACT_MAX_SIDE = 8192          # largest Gram matrix of the exact path (float64: 0.5 GB)
_EIGH_LOCK = threading.Lock()


def _eigh(gram):
    """torch.linalg.eigh behind a lock: torch loads its CUDA linear-algebra backend lazily on the first call, and that
    loader must not run in two threads at once ("lazy wrapper should be called at most once" when the workers of
    wot_b200.pipeline hit their first small local PCA together)."""
    import torch
    with _EIGH_LOCK:
        return torch.linalg.eigh(gram)


def compute_pca_gpu_exact(m1, m2, n_components, device=None):
    """Exact-solver path of util.py:240-255 on the GPU, for the shapes where scikit-learn's PCA(svd_solver='auto') does
    not run its randomized solver (few cells, or k close to the rank: `covariance_eigh` / `full`) and for
    rank-deficient input.  Everything stays in float64 on the device: centring, the Gram matrix of the SHORT side of
    the transposed matrix (library DGEMM through torch), its symmetric eigendecomposition (cuSOLVER through
    torch.linalg.eigh), the other side's vectors by one more product.  scikit-learn's sign convention
    (svd_flip(u_based_decision=False)) is applied, so components equal its LAPACK result up to roundoff wherever the
    spectrum is non-degenerate.  There is no CPU path: without a CUDA device this raises."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.WotB200Error("local PCA needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", _lib.context().device if device is None else int(device))
    a = torch.from_numpy(np.ascontiguousarray(_dense(m1), dtype=np.float64)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(_dense(m2), dtype=np.float64)).to(dev)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("m1 and m2 must be 2-D with the same number of genes")
    n1, n2, genes = a.shape[0], b.shape[0], a.shape[1]
    cells = n1 + n2
    k = min(int(n_components), cells)                        # util.py:247
    if k > min(cells, genes):
        raise ValueError("n_components=%d must be between 0 and min(n_samples, n_features)=%d" % (k, min(cells, genes)))
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    stacked = torch.cat([a, b], dim=0)                       # [cells, genes]
    gene_means = stacked.mean(dim=0)
    stacked = stacked - gene_means                           # util.py:244-246
    cell_means = stacked.mean(dim=1)                         # sklearn centres the transposed matrix by feature = cell
    xc = stacked - cell_means[:, None]                       # xc.T is the centred [genes, cells] matrix PCA factorises
    if min(cells, genes) > EXACT_MAX_SIDE:
        raise ValueError("exact local PCA: the short side (%d) exceeds %d" % (min(cells, genes), EXACT_MAX_SIDE))
    if cells <= genes:
        gram = xc @ xc.T                                     # [cells, cells] = V S^2 V^T
        lam, vec = _eigh(gram)
        lam, vec = lam.flip(0)[:k].clamp_min(0.0), vec.flip(1)[:, :k]
        comp = vec                                           # [cells, k]: right singular vectors of xc.T
    else:
        gram = xc.T @ xc                                     # [genes, genes] = U S^2 U^T
        lam, vec = _eigh(gram)
        lam, vec = lam.flip(0)[:k].clamp_min(0.0), vec.flip(1)[:, :k]
        comp = xc @ vec                                      # V = X^T U / S; |X^T u| = s, normalised directly so that
        nrm = comp.norm(dim=0)                               # directions of (numerically) zero singular values stay finite
        comp = comp / torch.where(nrm > 0, nrm, torch.ones_like(nrm))
    sv = lam.sqrt()
    top = comp.abs().argmax(dim=0)
    sign = torch.sign(comp[top, torch.arange(k, device=dev)])
    comp = comp * torch.where(sign == 0, torch.ones_like(sign), sign)
    stop.record()
    stop.synchronize()
    comp_h = comp.cpu().numpy()
    pca = LocalPCA(np.ascontiguousarray(comp_h.T), sv.cpu().numpy(), genes, cell_means.cpu().numpy(),
                   start.elapsed_time(stop))
    return comp_h[:n1], comp_h[n1:], pca, gene_means.cpu().numpy()


def compute_pca(m1, m2, n_components, backend="auto"):
    """Joint PCA of two cell populations, fitted on the TRANSPOSED, gene-mean-centred matrix.

    Returns (pca_1 [I, n], pca_2 [J, n], fitted PCA, gene means), like util.py:240-255.
    backend: 'auto' -- always on the GPU: the randomized solver (csrc/pca.cu) where scikit-learn's svd_solver='auto'
    would run its randomized solver, the exact solver (compute_pca_gpu_exact) for the small / near-full-rank shapes
    where scikit-learn itself switches to LAPACK and for rank-deficient input; 'gpu' = the randomized solver,
    'gpu_exact' = the exact one, 'sklearn' = the reference's own call (comparisons and tests only)."""
    if backend not in ("auto", "gpu", "gpu_exact", "sklearn"):
        raise ValueError("backend must be 'auto', 'gpu', 'gpu_exact' or 'sklearn'")
    n1, n2 = m1.shape[0], m2.shape[0]
    genes = m1.shape[1]
    k = min(int(n_components), n1 + n2)
    auto = backend == "auto"
    if auto:
        fits = k + N_OVERSAMPLES <= min(64, genes, n1 + n2)
        backend = "gpu" if fits and sklearn_solver_choice(genes, n1 + n2, k) == "randomized" else "gpu_exact"
    if backend == "sklearn":
        return compute_pca_sklearn(m1, m2, n_components)
    if backend == "gpu_exact":
        return compute_pca_gpu_exact(m1, m2, n_components)
    try:
        return compute_pca_gpu(m1, m2, n_components)
    except ValueError as exc:
        # numerically rank-deficient input (fewer independent cells or genes than k + 10 test vectors): the
        # Cholesky-QR of the range finder has no positive pivot; the exact solver has no such requirement (the
        # components that belong to zero singular values are arbitrary there, as they are in scikit-learn)
        if auto and "not positive definite" in str(exc):
            logging.getLogger("wot").warning("local PCA: rank-deficient input, using the exact GPU solver (%s)", exc)
            return compute_pca_gpu_exact(m1, m2, n_components)
        raise


def interpolate_with_ot(p0, p1, tmap, interp_frac, size):
    """Interpolated population of `size` cells between p0 and p1 at fraction `interp_frac`
    (reference: wot/ot/util.py:109-147): cell pairs (i, j) are drawn with probability
    tmap_ij / colsum_j^(1 - interp_frac) and mixed as p0[i] (1 - frac) + p1[j] frac.

    `tmap` is a wot_b200.tmap.ImplicitTransportMap (OTModel.compute_implicit_transport_map): the coupling is never
    materialised, the draw runs on the GPU in float64 (wotb_coupling_sample_host) and consumes the global NumPy
    random stream exactly like the reference's np.random.choice (one uniform sample per cell), so a seeded
    reference run and a seeded run of this function pick the same pairs."""
    from ..tmap import ImplicitTransportMap
    if not isinstance(tmap, ImplicitTransportMap):
        raise TypeError("interpolate_with_ot needs an ImplicitTransportMap (OTModel.compute_implicit_transport_map); "
                        "there is no CPU path")
    p0 = np.asarray(_dense(p0), dtype=np.float64)
    p1 = np.asarray(_dense(p1), dtype=np.float64)
    if p0.shape[1] != p1.shape[1]:
        raise ValueError("Unable to interpolate. Number of genes do not match")
    if p0.shape[0] != tmap.shape[0] or p1.shape[0] != tmap.shape[1]:
        raise ValueError("Unable to interpolate. Tmap size is {}, expected {}".format(tmap.shape, (len(p0), len(p1))))
    uniforms = np.random.random_sample(size)        # what np.random.choice(..., p=p, size=size) draws internally
    rows, cols = tmap.sample_pairs(interp_frac, uniforms)
    return p0[rows] * (1 - interp_frac) + p1[cols] * interp_frac
