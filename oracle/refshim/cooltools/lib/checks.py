"""Validation helpers: the reference only uses them as gates (raise_errors=True)."""


def is_compatible_viewframe(view_df, verify_cooler, check_sorting=False, raise_errors=False):
    sizes = verify_cooler.chromsizes
    ok = all(c in sizes.index and 0 <= s <= e <= sizes[c] for c, s, e in zip(view_df["chrom"], view_df["start"], view_df["end"]))
    if not ok and raise_errors:
        raise ValueError("view_df is out of the bounds of the cooler")
    return ok


def is_valid_expected(expected_df, contact_type, view_df, verify_cooler=None, expected_value_cols=(), raise_errors=False):
    need = ["region1", "region2"] + (["dist"] if contact_type == "cis" else []) + list(expected_value_cols)
    ok = all(c in expected_df.columns for c in need)
    if not ok and raise_errors:
        raise ValueError(f"expected is missing columns: {need}")
    return ok
