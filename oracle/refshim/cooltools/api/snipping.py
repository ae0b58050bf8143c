"""``cooltools.api.snipping.ExpectedSnipper.select`` (the only method the reference calls).

Documented behaviour: for ``region1 == region2`` returns
``LazyToeplitz(expected rows of (region1, region2)[expected_value_col].values)``
in table row order, i.e. ``exp[i, j] = E[|i - j|]``.  ``min_diag`` is only
applied by ``.snip`` (never called by coolpuppy).
"""
from ..numutils import LazyToeplitz


class ExpectedSnipper:
    def __init__(self, clr, expected, view_df=None, min_diag=2, expected_value_col="balanced.avg"):
        self.clr = clr
        self.expected = expected
        self.view_df = view_df.set_index("name") if "name" in view_df.columns else view_df
        self.min_diag = min_diag
        self.expected_value_col = expected_value_col

    def select(self, region1, region2):
        if region1 != region2:
            raise ValueError("ExpectedSnipper is implemented for cis contacts only")
        grp = self.expected.groupby(["region1", "region2"]).get_group((region1, region2))
        return LazyToeplitz(grp[self.expected_value_col].values)
