"""The little I/O the transport-map path touches (reference: wot/io/io.py).  Format parsing in general is
out of scope; when the reference package `wot` is installed its readers are used."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from ._anndata import AnnData, HAVE_ANNDATA


def check_file_extension(name, output_format):
    """io.py:475-478."""
    if not str(name).lower().endswith("." + output_format):
        name += "." + output_format
    return name


def check_output_format(output_format):
    """Raise for a format this installation cannot write BEFORE any transport map is computed."""
    if output_format not in ("h5ad", "loom", "txt", "npz"):
        raise ValueError("Unknown file format")
    if output_format == "loom" and not HAVE_ANNDATA:
        raise ImportError("writing .loom needs the anndata and loompy packages; use 'h5ad', 'txt' or 'npz'")


class TmapWriter:
    """write_dataset behind the caller's back: at most `depth` maps wait for the disk while the next ones are being
    solved (each waiting map pins its 0.2-3 GB block, hence the bound).  close() waits and re-raises a failed write."""

    def __init__(self, output_format, depth=2):
        from . import h5ad
        self.output_format = output_format
        self._async = h5ad.AsyncWriter(depth=depth)

    def write(self, ds, path):
        fmt = self.output_format
        self._async.submit(lambda: write_dataset(ds, path, output_format=fmt))

    def close(self, quiet=False):
        try:
            self._async.close()
        except BaseException:
            if not quiet:
                raise


def write_dataset(ds, path, output_format="txt"):
    """io.py:439-452, plus an 'npz' format usable without anndata/h5py."""
    path = check_file_extension(str(path), output_format)
    if output_format == "txt":
        pd.DataFrame(np.asarray(ds.X), index=ds.obs.index, columns=ds.var.index).to_csv(
            path, index_label="id", sep="\t", doublequote=False)
    elif output_format == "h5ad":
        ds.write(path)
    elif output_format == "loom":
        ds.write_loom(path)
    elif output_format == "npz":
        np.savez(path, X=np.asarray(ds.X), obs_index=np.asarray(ds.obs.index, dtype=str),
                 var_index=np.asarray(ds.var.index, dtype=str),
                 obs_columns=np.asarray(ds.obs.columns, dtype=str), obs_values=ds.obs.values)
    else:
        raise ValueError("Unknown file format")


def read_dataset(path):
    """Expression matrix reader: .h5ad via anndata, .txt/.tsv/.csv (cells x genes with an id column), .npz."""
    path = str(path)
    low = path.lower()
    if low.endswith(".h5ad"):
        if HAVE_ANNDATA:
            import anndata
            return anndata.read_h5ad(path)
        # the built-in reader handles the dense, uncompressed classic-format layout (what this package writes and
        # what anndata writes by default for a dense float matrix with float obs columns)
        from . import h5ad
        d = h5ad.read_h5ad(path)
        obs = pd.DataFrame({k: v for k, v in d["obs"].items()}, index=pd.Index(d["obs_index"].astype(str)))
        return AnnData(d["X"], obs, pd.DataFrame(index=pd.Index(d["var_index"].astype(str))))
    if low.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        obs = pd.DataFrame(index=pd.Index(z["obs_index"].astype(str)))
        if "obs_columns" in z.files:
            for k, name in enumerate(z["obs_columns"]):
                obs[str(name)] = z["obs_values"][:, k]
        return AnnData(z["X"], obs, pd.DataFrame(index=pd.Index(z["var_index"].astype(str))))
    df = pd.read_csv(path, index_col=0, engine="python", sep=None)
    return AnnData(df.values, pd.DataFrame(index=df.index.astype(str)), pd.DataFrame(index=df.columns.astype(str)))


def read_days_data_frame(path):
    return pd.read_csv(path, index_col="id", engine="python", sep=None, dtype={"day": np.float64})


def add_row_metadata_to_dataset(dataset, days=None, growth_rates=None, covariate=None):
    """io.py:526-545: join day / growth-rate / covariate tables on the cell id; growth defaults to 1."""
    if days is not None:
        if not os.path.exists(days):
            raise ValueError(days + " not found")
        dataset.obs = dataset.obs.join(read_days_data_frame(days))
    if growth_rates is not None:
        if not os.path.exists(growth_rates):
            raise ValueError(growth_rates + " not found")
        dataset.obs = dataset.obs.join(pd.read_csv(growth_rates, index_col="id", engine="python", sep=None))
    else:
        dataset.obs["cell_growth_rate"] = 1.0
    if covariate is not None:
        if not os.path.exists(covariate):
            raise ValueError(covariate + " not found")
        dataset.obs = dataset.obs.join(pd.read_csv(covariate, index_col="id", engine="python", sep=None))


def read_day_pairs(day_pairs):
    """io.py:548-556: a file, or an inline 't0,t1;...' string."""
    if os.path.isfile(day_pairs):
        return pd.read_csv(day_pairs, engine="python", sep=None)
    import io
    return pd.read_csv(io.StringIO(day_pairs), sep=",", lineterminator=";")


def filter_adata(adata, obs_filter=None, var_filter=None):
    """io.py:499-518 for id lists / boolean fields (set files are read as one id per line)."""
    def ids(spec):
        if os.path.exists(spec):
            with open(spec) as fh:
                return [ln.strip().split("\t")[0] for ln in fh if ln.strip()]
        return spec.split(",")

    if obs_filter is not None:
        sel = ids(obs_filter)
        if len(sel) == 1 and sel[0] in adata.obs:
            adata = adata[(adata.obs[sel[0]] == True).values].copy()  # noqa: E712
        else:
            adata = adata[adata.obs.index.isin(sel)].copy()
    if var_filter is not None:
        sel = ids(var_filter)
        if len(sel) == 1 and sel[0] in adata.var:
            adata = adata[:, np.asarray(adata.var[sel[0]], dtype=bool)].copy()
        else:
            adata = adata[:, adata.var.index.isin(sel)].copy()
    return adata
