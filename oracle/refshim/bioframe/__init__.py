"""Stand-in for the few ``bioframe`` (>=0.3.3) calls on the reference path.

* ``make_viewframe(regions, check_bounds=chromsizes)``: 3-column input gets
  ``name = chrom`` (bioframe's default ``name_style=None``), 4-column input
  keeps its names; bounds are validated.
* ``read_table(path, schema="bed"|"bed3"|"bed4"|"bedpe"|...)``: headerless TSV
  with the standard UCSC column names, truncated to the columns present.
* ``sort_bedframe(df, view_df)``: sort by view order of ``chrom`` then
  start/end; rows whose chrom is not in the view go last.
* ``expand(df, scale=...)``: intervals grown about their midpoints (only reached with ``rescale_flank``).
"""
import numpy as np
import pandas as pd

SCHEMAS = {
    "bed3": ["chrom", "start", "end"],
    "bed4": ["chrom", "start", "end", "name"],
    "bed": ["chrom", "start", "end", "name", "score", "strand"],
    "bed6": ["chrom", "start", "end", "name", "score", "strand"],
    "bedpe": ["chrom1", "start1", "end1", "chrom2", "start2", "end2", "name", "score", "strand1", "strand2"],
}


def read_table(filepath_or, schema=None, schema_is_strict=False, **kwargs):
    kwargs.setdefault("sep", "\t")
    kwargs.setdefault("header", None)
    kwargs.setdefault("comment", "#")
    df = pd.read_csv(filepath_or, **kwargs)
    if schema is not None:
        names = SCHEMAS[schema]
        n = min(len(names), df.shape[1])
        df = df.iloc[:, : max(n, df.shape[1])]
        cols = list(names[:n]) + [f"col{i}" for i in range(n, df.shape[1])]
        df.columns = cols
    return df


def make_viewframe(regions, check_bounds=None, name_style=None, view_name_col="name", cols=None):
    ck, sk, ek = ("chrom", "start", "end") if cols is None else cols
    if isinstance(regions, pd.DataFrame):
        view = regions.copy()
        if view.shape[1] >= 3 and ck not in view.columns:
            view.columns = [ck, sk, ek, view_name_col][: view.shape[1]]
        if view_name_col not in view.columns:
            if name_style == "ucsc":
                view[view_name_col] = [f"{c}:{s}-{e}" for c, s, e in zip(view[ck], view[sk], view[ek])]
            else:
                view[view_name_col] = view[ck].values
        view = view[[ck, sk, ek, view_name_col]].reset_index(drop=True)
    elif isinstance(regions, (dict, pd.Series)):
        items = dict(regions)
        view = pd.DataFrame(
            {ck: list(items), sk: 0, ek: [int(v) for v in items.values()], view_name_col: list(items)}
        )
    else:
        raise ValueError("unsupported view specification")
    view[ck] = view[ck].astype(str)
    view[view_name_col] = view[view_name_col].astype(str)
    if view[view_name_col].duplicated().any():
        raise ValueError("view names must be unique")
    if check_bounds is not None:
        sizes = dict(check_bounds)
        for c, s, e in zip(view[ck], view[sk], view[ek]):
            if c not in sizes or s < 0 or e > sizes[c]:
                raise ValueError(f"view region {c}:{s}-{e} out of bounds")
    return view


def sort_bedframe(df, view_df=None, reset_index=True, df_view_col=None, view_name_col="name", cols=None):
    ck, sk, ek = ("chrom", "start", "end") if cols is None else cols
    out = df.copy()
    if view_df is not None:
        order = {c: i for i, c in enumerate(pd.unique(view_df[ck]))}
        rank = out[ck].map(lambda c: order.get(c, len(order)))
    else:
        rank = out[ck]
    out = out.assign(_rank=rank.values).sort_values(["_rank", sk, ek], kind="stable").drop(columns="_rank")
    return out.reset_index(drop=True) if reset_index else out


def expand(df, pad=None, scale=None, side="both", cols=None):
    """``bioframe.expand`` (bioframe/ops.py) with ``scale``: every interval grows about its midpoint to ``scale`` times
    its length -- ``pads = 0.5 * (scale - 1) * (end - start)``, ``start - pads`` / ``end + pads``, rounded (half to
    even, ``DataFrame.round``) and cast back to the columns' integer dtypes.  ``pad`` (additive) is not used by the
    reference.  Restated from bioframe's source as remembered; no artefact in the reference tree pins it."""
    ck, sk, ek = ("chrom", "start", "end") if cols is None else cols
    if (scale is None) == (pad is None):
        raise ValueError("exactly one of pad or scale is needed")
    if pad is not None or side != "both":
        raise NotImplementedError("only expand(scale=..., side='both') is on the reference path")
    if scale < 0:
        raise ValueError("multiplicative scale must be >= 0")
    out = df.copy()
    pads = 0.5 * (scale - 1) * (df[ek].values - df[sk].values)
    types = df.dtypes[[sk, ek]]
    out[sk] = df[sk].values - pads
    out[ek] = df[ek].values + pads
    out[[sk, ek]] = out[[sk, ek]].round()
    out[[sk, ek]] = out[[sk, ek]].astype(types)
    return out
