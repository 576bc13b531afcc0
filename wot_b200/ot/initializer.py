"""Configuration plumbing of the transport-map path (reference: wot/ot/initializer.py)."""
from __future__ import annotations

import pandas as pd

from .. import io as _io

_PAIR_FIELDS = ("epsilon", "lambda1", "lambda2")  # the only keys that may vary per day-pair, initializer.py:176


def initialize_ot_model(matrix, **kwargs):
    """File -> OTModel (initializer.py:8-39): read the matrix, join days / growth rates / covariates."""
    from .ot_model import OTModel
    ds = _io.read_dataset(matrix)
    if kwargs.pop("transpose", False):
        ds = ds.T
    _io.add_row_metadata_to_dataset(dataset=ds, days=kwargs.pop("cell_days", None),
                                    growth_rates=kwargs.pop("cell_growth_rates", None),
                                    covariate=kwargs.pop("covariate", None))
    return OTModel(ds, **kwargs)


def parse_parameter_file(path):
    """Two-column file of parameter name and value (initializer.py:126-132)."""
    table = pd.read_csv(path, engine="python", sep=None, header=None)
    return {table.iloc[k, 0]: table.iloc[k, 1] for k in range(len(table))}


def parse_configuration(config):
    """None | path or inline string | DataFrame with column t, or t0 and t1 (initializer.py:42-81)."""
    if config is None:
        return None
    if isinstance(config, str):
        return parse_configuration(_io.read_day_pairs(config))
    if isinstance(config, pd.DataFrame):
        if "t" in config.columns:
            return parse_per_timepoint_configuration(config)
        if "t0" in config.columns and "t1" in config.columns:
            return parse_per_timepair_configuration(config)
        raise ValueError("Configuration must have at least a column 't' or two columns 't0' and 't1'")
    if isinstance(config, dict):
        raise ValueError("Not implemented")
    raise ValueError("Unrecognized argument type for configuration. Use DataFrame, dict, str or None")


def parse_per_timepoint_configuration(config):
    """Per-timepoint values become per-pair values by averaging neighbours (initializer.py:84-123)."""
    if isinstance(config, dict):
        raise ValueError("Not implemented")
    if not isinstance(config, pd.DataFrame):
        raise ValueError("Unrecognized argument type for per-timepoint configuration. Use DataFrame, str or None")
    if "t" not in config.columns:
        raise ValueError("Invalid per-timepoint configuration : must have column t")
    casts = {c: float for c in ("t",) + _PAIR_FIELDS if c in config.columns}
    table = config.sort_values(by="t").astype(casts)
    fields = [c for c in table.columns if c in _PAIR_FIELDS]
    pairs = {}
    for k in range(len(table) - 1):
        lo, hi = table.iloc[k], table.iloc[k + 1]
        pairs[(lo["t"], hi["t"])] = {c: (lo[c] + hi[c]) / 2 for c in fields}
    return pairs


def parse_per_timepair_configuration(config):
    """DataFrame with t0, t1 (+ epsilon/lambda1/lambda2) or dict keyed by (t0, t1) (initializer.py:135-185)."""
    if isinstance(config, pd.DataFrame):
        if "t0" not in config.columns or "t1" not in config.columns:
            raise ValueError("Invalid per-timepair configuration : must have columns t0 and t1")
        as_dict = {}
        for k in range(len(config)):
            row = config.loc[k].to_dict()
            as_dict[(row.pop("t0"), row.pop("t1"))] = row
        return parse_per_timepair_configuration(as_dict)
    if isinstance(config, dict):
        try:
            ok = all(isinstance(z, (int, float)) for t0, t1 in config for z in (t0, t1))
        except Exception:
            raise ValueError("Dictionnary keys for config must be pairs")
        if not ok:
            raise ValueError("Dictionnary keys for config must be pairs")
        for key, val in config.items():
            if not isinstance(val, dict):
                raise ValueError("Dictionnary values for config must be dictionnaries")
            config[key] = {c: val[c] for c in _PAIR_FIELDS if c in val}
        return config
    raise ValueError("Unrecognized argument type for config. Use DataFrame or dict")
