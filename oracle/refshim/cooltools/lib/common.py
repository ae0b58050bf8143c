import pandas as pd


def make_cooler_view(clr, ucsc_names=False):
    """One view row per chromosome: (chrom, 0, length, name=chrom)."""
    names = list(clr.chromnames)
    df = pd.DataFrame(
        {"chrom": names, "start": 0, "end": [int(clr.chromsizes[c]) for c in names], "name": names}
    )
    if ucsc_names:
        df["name"] = [f"{c}:{s}-{e}" for c, s, e in zip(df.chrom, df.start, df.end)]
    return df
