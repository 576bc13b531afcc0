"""TEST INFRASTRUCTURE ONLY (see oracle/wot_oracle.py): float64 NumPy restatement of the local PCA the reference
runs before the cost (wot/ot/util.py:240-255: sklearn PCA(n_components, random_state=58951).fit(x.T) on the
gene-mean-centred matrix), written the way the CUDA path computes it.

The arithmetic lives in scikit-learn, which is not part of /root/reference (setup.py:8-10 leaves it unpinned; the
Terra image pins scikit-learn==0.22.2.post1).  Restated here from the published algorithm as implemented by
sklearn 1.9 (sklearn/decomposition/_pca.py::_fit_truncated, sklearn/utils/extmath.py::randomized_svd and
randomized_range_finder; Halko, Martinsson, Tropp 2011, algorithms 4.3/4.4/5.1):
  X_c = x.T - mean over genes (per cell);  n_random = k + 10;  n_iter = 7 if k < 0.1 min(shape) else 4;
  M = X_c, transposed when it has fewer rows than columns;  Q0 = RandomState(58951).normal(size=(M.shape[1], n_random));
  n_iter times: Q = normalise(M Q); Q = normalise(M^T Q);  Q = qr(M Q);  B = Q^T M;  svd(B) -> components, singular values.
sklearn normalises with a pivoted LU between power iterations; any normaliser that keeps the span gives the same
range (and therefore the same components up to sign and roundoff), so this restatement and the CUDA path use
Cholesky-QR.  PINNED twice: tests/test_oracle.py::test_pca_restatement_matches_reference_golden against fixtures produced by
the unmodified reference's wot.ot.compute_pca (tests/golden/pca_randomized.npz, make_golden.py pca_cases), and
::test_pca_restatement_matches_sklearn against the installed scikit-learn executing the reference's own call
(cost-matrix level, sign-invariant)."""
import numpy as np


def chol_qr2(y):
    """Orthonormal basis of span(y) by two rounds of Cholesky-QR (what the CUDA path does)."""
    for _ in range(2):
        r = np.linalg.cholesky(y.T @ y).T
        y = np.linalg.solve(r.T, y.T).T
    return y


def solver_choice(n_samples, n_features, k):
    """sklearn 1.9 PCA(svd_solver='auto') policy (_pca.py::_fit): which solver fit(x.T) ends up in."""
    if n_features <= 1000 and n_samples >= 10 * n_features:
        return "covariance_eigh"
    if max(n_samples, n_features) <= 500:
        return "full"
    if 1 <= k < 0.8 * min(n_samples, n_features):
        return "randomized"
    return "full"


def compute_pca(m1, m2, n_components, seed=58951, n_oversamples=10):
    """(pca_1 [I, k], pca_2 [J, k], singular_values [k], gene_means): randomized path of util.py:240-255."""
    x = np.vstack([np.asarray(m1, dtype=np.float64), np.asarray(m2, dtype=np.float64)])   # cells x genes
    gene_means = x.mean(axis=0)                                                           # :245
    x = x - gene_means                                                                    # :246
    k = min(n_components, x.shape[0])                                                     # :247
    a = x - x.mean(axis=1, keepdims=True)      # PCA.fit(x.T) centres every feature (= cell) over the samples (= genes)
    n_cells, n_genes = a.shape
    m = a.T                                    # what sklearn calls X_centered: genes x cells
    size = k + n_oversamples
    n_iter = 7 if k < 0.1 * min(m.shape) else 4
    transpose = m.shape[0] < m.shape[1]
    if transpose:
        m = m.T                                # cells x genes
    rng = np.random.RandomState(seed)
    q = rng.normal(size=(m.shape[1], size))
    for _ in range(n_iter):
        q = chol_qr2(m @ q)
        q = chol_qr2(m.T @ q)
    q = chol_qr2(m @ q)
    b = q.T @ m
    uhat, s, vt = np.linalg.svd(b, full_matrices=False)
    if transpose:
        comp = (q @ uhat)[:, :k]               # cells x k
    else:
        comp = vt[:k].T
    n1 = np.asarray(m1).shape[0]
    return comp[:n1], comp[n1:], s[:k], gene_means
