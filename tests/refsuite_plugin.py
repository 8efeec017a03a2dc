"""pytest plugin (loaded with `-p refsuite_plugin`) that runs the REFERENCE's own test files
against this repo's implementation: the reference package is imported from /root/reference/src -- on the GPU
box from the vendored, git-ignored baseline/_ref (oracle/make_ref.py) -- with a stub matplotlib, and its six
hot-path operations plus the two `_intidxs` helpers are replaced by
`bioframe_b200.ops`, so `bioframe.overlap(...)` in a reference test calls our glue -> our engine.
Derived reference ops (subtract, setdiff, count_overlaps, trim, ...) then run on top of ours too.
REFSUITE_ENGINE=oracle: the engine is the oracle-backed test double (tests/oracle_engine.py): that checks the
pandas glue.  REFSUITE_ENGINE=cuda: the CUDA engine (fails loudly without a GPU).  Unset: cuda when available."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _install():
    sys.dont_write_bytecode = True
    if os.environ.get("REFSUITE_INFER_STRING", "0") == "0":
        # pandas 3 infers str columns as StringDtype; three reference tests expect object dtype
        import pandas as pd

        pd.set_option("future.infer_string", False)
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from conftest import import_reference

    bioframe = import_reference()
    import bioframe.ops as ref_ops

    import bioframe_b200.core.arrops as our_arrops
    import bioframe_b200.ops as our_ops

    want = os.environ.get("REFSUITE_ENGINE", "")
    use_gpu = False
    try:
        import torch

        use_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if want == "cuda" and not use_gpu:
        raise RuntimeError("REFSUITE_ENGINE=cuda but no CUDA device")
    if want == "oracle":
        use_gpu = False
    if not use_gpu:
        import oracle_engine

        our_ops._engine = oracle_engine
        our_arrops._engine = oracle_engine
    for name in ("overlap", "closest", "cluster", "merge", "complement", "coverage", "count_overlaps", "setdiff",
                 "subtract", "assign_view", "trim", "_overlap_intidxs", "_closest_intidxs"):
        setattr(ref_ops, name, getattr(our_ops, name))
        if hasattr(bioframe, name):
            setattr(bioframe, name, getattr(our_ops, name))
    import bioframe.extras as ref_extras

    import bioframe_b200.extras as our_extras

    for name in ("pair_by_distance", "mark_runs", "merge_runs"):
        setattr(ref_extras, name, getattr(our_extras, name))
        if hasattr(bioframe, name):
            setattr(bioframe, name, getattr(our_extras, name))
    print(f"[refsuite] reference ops patched with bioframe_b200 ({'CUDA engine' if use_gpu else 'oracle test double'})")


_install()
