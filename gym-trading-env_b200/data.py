"""Host-side data staging for the batched trading env (numpy / pandas only, runs once).

Mirrors what the reference does once per DataFrame in ``TradingEnv._set_df``
(`/root/reference/src/gym_trading_env/environments.py:128-143`):

* feature columns  = every column whose name contains ``"feature"``, in DataFrame order (`:130`);
* the float32 feature matrix is produced by *numpy's own* float64→float32 cast (`:141`), so the
  device copy is bit-identical to the reference's ``_obs_array`` by construction;
* the price vector is the float64 ``close`` column (`:143`).

Also holds the synthetic GBM OHLCV generator every BASELINE.json config is quoted on
(SURVEY.md §8(d)).  Nothing here touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd


@dataclass
class SeriesArrays:
    """One market series, staged the way `_set_df` stages it (environments.py:128-143)."""

    features: np.ndarray      # float32 [T, F_static]  (static "feature*" columns, df order)
    price: np.ndarray         # float64 [T]            ("close")
    feature_names: list
    info: dict                # name -> float64 [T] for numeric non-feature columns (open/high/low/close/volume…)
    index: np.ndarray | None  # datetime64 index values (infos["date"]); may be None

    @property
    def length(self) -> int:
        return int(self.price.shape[0])


def frame_to_arrays(df: pd.DataFrame) -> SeriesArrays:
    """DataFrame -> arrays with the reference's column rules (environments.py:130-143)."""
    if "close" not in df.columns:
        raise ValueError("the DataFrame must contain a 'close' column (environments.py:143)")
    feature_names = [c for c in df.columns if "feature" in c]                      # :130
    feats = np.array(df[feature_names], dtype=np.float32) if feature_names else \
        np.zeros((len(df), 0), dtype=np.float32)                                   # :141
    feats = np.ascontiguousarray(feats.reshape(len(df), len(feature_names)))
    price = np.ascontiguousarray(np.array(df["close"], dtype=np.float64))          # :143
    info = {}
    for c in df.columns:
        if c in feature_names:
            continue
        col = df[c]
        if pd.api.types.is_numeric_dtype(col.dtype):
            info[c] = np.ascontiguousarray(np.array(col, dtype=np.float64))
    index = df.index.values if isinstance(df.index, pd.DatetimeIndex) else None
    return SeriesArrays(feats, price, feature_names, info, index)


def load_frame(path: str) -> pd.DataFrame:
    """One dataset file -> DataFrame.  ``*.pkl`` / ``*.pickle`` as the reference reads them (``pd.read_pickle``,
    environments.py:391 — what its downloader writes, downloader.py:63-70); ``*.csv`` / ``*.csv.gz`` the way the
    reference's examples load the bundled file (examples/example_environnement.py:11-14): a ``date`` column (or
    ``timestamp`` / ``datetime``) becomes the parsed, sorted DatetimeIndex; ``*.parquet`` via pandas."""
    low = path.lower()
    if low.endswith((".pkl", ".pickle")):
        return pd.read_pickle(path)
    if low.endswith((".csv", ".csv.gz")):
        df = pd.read_csv(path)
        for c in ("date", "timestamp", "datetime", "Date"):
            if c in df.columns:
                df[c] = pd.to_datetime(df[c])
                df = df.set_index(c).sort_index()
                df.index.name = "date"
                break
        return df
    if low.endswith(".parquet"):
        return pd.read_parquet(path)
    raise ValueError(f"unsupported dataset file type: {path} (expected .pkl, .csv or .parquet)")


def reconcile_series(series, names=None):
    """Make a list of staged datasets stackable into one device table: same static feature columns in the SAME order.

    The reference builds one observation space from whatever frame it loaded first and simply breaks (shape mismatch in
    ``_get_obs``) when a later frame has other feature columns.  Here the first dataset's feature names are the schema;
    every other dataset is re-ordered BY NAME to match it, and a dataset that lacks one of them (or carries an extra
    one) is refused with a message that names it.  Lengths may differ (ragged: each dataset keeps its own T); the info
    columns (``data_*`` of the info dict) are reduced to those every dataset has.  Non-numeric columns never reach
    this point (``frame_to_arrays`` keeps numeric columns only)."""
    if not series:
        return series
    names = names or [f"dataset {k}" for k in range(len(series))]
    schema = list(series[0].feature_names)
    out = []
    for srs, nm in zip(series, names):
        have = list(srs.feature_names)
        if have != schema:
            missing, extra = [c for c in schema if c not in have], [c for c in have if c not in schema]
            if missing or extra or len(have) != len(schema):
                raise ValueError(f"{nm}: feature columns differ from the first dataset's — missing {missing}, "
                                 f"unexpected {extra} (expected {schema})")
            order = [have.index(c) for c in schema]
            srs = SeriesArrays(np.ascontiguousarray(srs.features[:, order]), srs.price, schema, srs.info, srs.index)
        out.append(srs)
    common = [c for c in out[0].info if all(c in s.info for s in out)]
    return [SeriesArrays(s.features, s.price, s.feature_names, {c: s.info[c] for c in common}, s.index) for s in out]


def make_gbm_ohlcv(T: int = 100_000, seed: int = 0, sigma: float = 0.002) -> pd.DataFrame:
    """Synthetic GBM OHLCV frame with 8 static ``feature_*`` columns.

    log-returns r_t ~ N(0, sigma); close = 100*exp(cumsum r); open/high/low jittered around it;
    volume ~ U(1,2); hourly DatetimeIndex from 2000-01-01.  T+168 rows are generated, NaN rows
    from the rolling/pct_change features are dropped, and exactly the last T rows are kept.
    Feature recipe = `/root/reference/examples/example_environnement.py:18-22` plus three
    longer-horizon columns so that F_static = 8 as BASELINE.json's configs state.
    """
    rng = np.random.default_rng(seed)
    n = T + 168
    r = rng.normal(0.0, sigma, size=n)
    close = 100.0 * np.exp(np.cumsum(r))
    open_ = close * (1.0 + rng.normal(0.0, 1e-3, size=n))
    high = np.maximum(open_, close) * (1.0 + np.abs(rng.normal(0.0, 1e-3, size=n)))
    low = np.minimum(open_, close) * (1.0 - np.abs(rng.normal(0.0, 1e-3, size=n)))
    volume = rng.uniform(1.0, 2.0, size=n)
    idx = pd.date_range("2000-01-01", periods=n, freq="h")
    df = pd.DataFrame({"open": open_, "high": high, "low": low, "close": close, "volume": volume}, index=idx)
    df["feature_close"] = df["close"].pct_change()
    df["feature_open"] = df["open"] / df["close"]
    df["feature_high"] = df["high"] / df["close"]
    df["feature_low"] = df["low"] / df["close"]
    df["feature_volume"] = df["volume"] / df["volume"].rolling(7 * 24).max()
    df["feature_ret_8"] = df["close"].pct_change(8)
    df["feature_ret_64"] = df["close"].pct_change(64)
    df["feature_vol_64"] = df["close"].pct_change().rolling(64).std()
    df = df.dropna()
    df = df.iloc[-T:].copy()
    assert len(df) == T, (len(df), T)
    return df


def make_gbm_arrays(T: int, seed: int = 0, sigma: float = 0.002, n_features: int = 8):
    """Fast array-only GBM series for large benchmark tables (no pandas; C4 uses 32 x 1M rows).

    Same price process as :func:`make_gbm_ohlcv`; features are cheap functions of the same
    series (returns over several horizons), float32-cast by numpy as `_set_df` does (:141).
    Used only where building a DataFrame per series would dominate set-up time.
    """
    rng = np.random.default_rng(seed)
    r = rng.normal(0.0, sigma, size=T + 64)
    close = 100.0 * np.exp(np.cumsum(r))
    feats = np.empty((T, n_features), dtype=np.float64)
    horizons = [1, 2, 4, 8, 16, 32, 64]
    for j in range(n_features):
        h = horizons[j % len(horizons)]
        feats[:, j] = close[64:] / close[64 - h:T + 64 - h] - 1.0
    return np.ascontiguousarray(feats.astype(np.float32)), np.ascontiguousarray(close[64:])


def window_table_classes(row_bytes: int):
    """Alignment classes c = ((r0*row_bytes) >> 2) & 3 a window start row r0 can fall in."""
    return sorted({((r * row_bytes) >> 2) & 3 for r in range(4)})


def build_window_tables(features: np.ndarray, n_dyn: int):
    """Host-side layout of the 16-byte-aligned window tables used by the vector / TMA gathers.

    features: float32 [n_datasets, t_stride, n_static].  Returns ``(tables, shifts, ds_stride)``:
    ``tables[c]`` is a uint8 buffer (to be placed at a 16-byte-aligned device address) holding every
    dataset's rows in the reference's own ``_obs_array`` layout ``[t, n_static+n_dyn]`` (dynamic
    columns zero, environments.py:135-141), starting ``shifts[c] = (16 - 4c) % 16`` bytes into the
    buffer, datasets ``ds_stride`` bytes apart; a window whose first row r0 has
    ``(r0*row_bytes) % 16 == 4c`` then starts on a 16-byte boundary of copy c.
    """
    n_ds, t_stride, ns = features.shape
    F = ns + n_dyn
    row_bytes = 4 * F
    rows = np.zeros((n_ds, t_stride, F), dtype=np.float32)
    rows[:, :, :ns] = features
    raw = rows.reshape(n_ds, -1).view(np.uint8)
    ds_stride = ((t_stride * row_bytes + 15) // 16) * 16
    tables, shifts = {}, {}
    for c in window_table_classes(row_bytes):
        shift = (16 - 4 * c) % 16
        host = np.zeros(n_ds * ds_stride + 16, dtype=np.uint8)
        for k in range(n_ds):
            o = shift + k * ds_stride
            host[o:o + raw.shape[1]] = raw[k]
        tables[c], shifts[c] = host, shift
    return tables, shifts, ds_stride
