"""The checker of the full-size parity tests: the REAL reference (`bioframe.ops._overlap_intidxs`,
`_closest_intidxs`, `cluster`, `merge`, `coverage` from /root/reference or the vendored baseline/_ref) when it
is importable, else the numpy oracle port.  Group codes are handed to the reference as an integer `chrom`
column: its group loop walks `set.union(...)` of the keys (ops.py:289-301), which for small ints is ascending
order -- the group-rank order of the C ABI -- so its row order is reproducible (for strings it follows the
process's hash seed)."""
import numpy as np

from conftest import import_reference, reference_paths
from oracle import ops_oracle as oo


def have_reference():
    return reference_paths()[0] is not None


def _frame(g, s, e):
    import pandas as pd

    g = np.zeros(len(s), dtype=np.int64) if g is None else np.asarray(g).astype(np.int64)
    return pd.DataFrame({"chrom": g, "start": np.asarray(s), "end": np.asarray(e)})


def overlap(g1, s1, e1, g2, s2, e2, n_groups, how):
    """-> (events1, events2, kind)"""
    if have_reference():
        bf = import_reference()
        a, b = bf.ops._overlap_intidxs(_frame(g1, s1, e1), _frame(g2, s2, e2), how=how)
        return np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64), "reference"
    a, b = oo.overlap_intidxs_coded(g1, s1, e1, g2, s2, e2, n_groups, how=how)
    return a, b, "port"


def closest(g1, s1, e1, g2, s2, e2, n_groups, **kw):
    if have_reference():
        bf = import_reference()
        a, b = bf.ops._closest_intidxs(_frame(g1, s1, e1), _frame(g2, s2, e2), **kw)
        return np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64), "reference"
    a, b = oo.closest_intidxs_coded(g1, s1, e1, g2, s2, e2, n_groups, **kw)
    return a, b, "port"


def cluster(g, s, e, n_groups, min_dist=0):
    """-> (ids, cluster_start per row, cluster_end per row, merged table (chrom, start, end, n), kind)"""
    if have_reference():
        bf = import_reference()
        import pandas as pd

        # cluster / merge index their group keys like tuples (ops.py:666): strings whose sorted order is code order
        names = np.array([f"c{c:04d}" for c in range(n_groups)], dtype=object)
        g = np.zeros(len(s), dtype=np.int64) if g is None else np.asarray(g)
        df = pd.DataFrame({"chrom": names[g], "start": np.asarray(s), "end": np.asarray(e)})
        c = bf.cluster(df, min_dist=min_dist)
        m = bf.merge(df, min_dist=min_dist)
        table = (m["chrom"].str.slice(1).astype(np.int64).to_numpy(),
                 *(m[k].to_numpy().astype(np.int64) for k in ("start", "end", "n_intervals")))
        return (c["cluster"].to_numpy().astype(np.int64), c["cluster_start"].to_numpy().astype(np.int64),
                c["cluster_end"].to_numpy().astype(np.int64), table, "reference")
    ids, cs, ce, cn = oo.cluster_coded(g, s, e, n_groups, min_dist)
    cg = oo.cluster_coded.last_groups
    return ids, cs[ids], ce[ids], (cg, cs, ce, cn), "port"


def coverage(g1, s1, e1, g2, s2, e2, n_groups):
    if have_reference():
        bf = import_reference()
        import pandas as pd

        names = np.array([f"c{c:04d}" for c in range(n_groups)], dtype=object)  # merge() wants string chroms

        def named(g, s, e):
            g = np.zeros(len(s), dtype=np.int64) if g is None else np.asarray(g)
            return pd.DataFrame({"chrom": names[g], "start": np.asarray(s), "end": np.asarray(e)})

        out = bf.coverage(named(g1, s1, e1), named(g2, s2, e2))
        return out["coverage"].to_numpy().astype(np.int64), "reference"
    return oo.coverage_coded(g1, s1, e1, g2, s2, e2, n_groups), "port"
