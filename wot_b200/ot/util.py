"""Local PCA, the producer of the hot path's inputs (reference: wot/ot/util.py:240-255).  Kept on the
CPU / scikit-learn so the GPU path and the reference see identical coordinates."""
from __future__ import annotations

import numpy as np
import scipy.sparse
import sklearn.decomposition


def compute_pca(m1, m2, n_components):
    """Joint PCA of two cell populations, fitted on the TRANSPOSED, gene-mean-centred matrix.

    Returns (pca_1 [I, n], pca_2 [J, n], fitted PCA, gene means), like util.py:240-255.
    """
    dense = [m.toarray() if scipy.sparse.isspmatrix(m) else np.asarray(m) for m in (m1, m2)]
    stacked = np.vstack(dense)
    gene_means = stacked.mean(axis=0)
    stacked = stacked - gene_means
    n_components = min(n_components, stacked.shape[0])  # cannot exceed the number of cells
    pca = sklearn.decomposition.PCA(n_components=n_components, random_state=58951)
    pca.fit(stacked.T)
    loadings = pca.components_.T
    n1 = dense[0].shape[0]
    return loadings[:n1], loadings[n1:n1 + dense[1].shape[0]], pca, gene_means
