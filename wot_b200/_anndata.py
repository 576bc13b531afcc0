"""AnnData when the `anndata` package is importable, otherwise the minimal stand-in the transport-map
path needs (X, obs, var, shape, boolean row masks, copy, write).  The reference constructs its result as
anndata.AnnData(tmap, obs, var) (ot_model.py:326); both classes accept that call."""
from __future__ import annotations

import numpy as np
import pandas as pd

try:  # pragma: no cover - depends on the environment
    import anndata as _anndata
    AnnData = _anndata.AnnData
    HAVE_ANNDATA = True
except Exception:  # anndata (and h5py) are not installed in the build image
    HAVE_ANNDATA = False

    class AnnData:  # type: ignore[no-redef]
        def __init__(self, X, obs=None, var=None):
            self.X = X
            n, m = X.shape
            self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n).astype(str))
            self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(m).astype(str))

        @property
        def shape(self):
            return self.X.shape

        @property
        def T(self):
            """Transposed view: cells <-> genes (initialize_ot_model(transpose=True), wot/ot/initializer.py)."""
            return AnnData(self.X.T, self.var, self.obs)

        def copy(self):
            return AnnData(self.X.copy(), self.obs.copy(), self.var.copy())

        def __getitem__(self, key):
            rows, cols = key if isinstance(key, tuple) else (key, slice(None))
            if isinstance(rows, pd.Series):
                rows = rows.values
            rows = np.asarray(rows) if not isinstance(rows, slice) else rows
            X = self.X[rows]
            if isinstance(rows, slice):
                obs = self.obs.iloc[rows]
            else:
                obs = self.obs[rows] if rows.dtype == bool else self.obs.iloc[rows]
            var = self.var
            if not (isinstance(cols, slice) and cols == slice(None)):
                if isinstance(cols, pd.Series):
                    cols = cols.values
                cols = np.asarray(cols)
                X = X[:, cols]
                var = self.var[cols] if cols.dtype == bool else self.var.iloc[cols]
            return AnnData(X, obs, var)

        def write(self, path):
            """.h5ad in the layout anndata writes and TransportMapModel.from_directory reads
            (wot/tmap/transport_map_model.py:709-721), through the built-in HDF5 writer (wot_b200/h5ad.py)."""
            from . import h5ad
            h5ad.write_anndata(path, self)

        def write_loom(self, path):
            raise ImportError("writing .loom needs the anndata and loompy packages; use 'h5ad', 'txt' or 'npz'")
