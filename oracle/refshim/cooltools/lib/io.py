"""read_viewframe_from_file / read_expected_from_file as used by the reference's tests."""
import pandas as pd
import bioframe


def read_viewframe_from_file(view_fname, verify_cooler=None, check_sorting=False):
    df = pd.read_csv(view_fname, sep="\t", header=None, comment="#")
    df = df.iloc[:, :4]
    df.columns = ["chrom", "start", "end", "name"][: df.shape[1]]
    return bioframe.make_viewframe(df, check_bounds=None if verify_cooler is None else verify_cooler.chromsizes)


def read_expected_from_file(fname, contact_type="cis", expected_value_cols=("count.avg", "balanced.avg"), verify_view=None, verify_cooler=None, raise_errors=True):
    df = pd.read_csv(fname, sep="\t", dtype={"region1": str, "region2": str})
    return df
