"""GPU BED ingest (bf_bed_*) against pandas.read_csv, which is what the reference's read_table calls
(src/bioframe/io/fileops.py:42-84)."""
import io

import numpy as np
import pandas as pd
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def object_strings():
    """object-dtype strings on both sides (pandas 3 would infer its string dtype for names and columns)"""
    old = pd.get_option("future.infer_string")
    pd.set_option("future.infer_string", False)
    yield
    pd.set_option("future.infer_string", old)


def _reference_frame(text):
    return pd.read_csv(io.BytesIO(text), sep="\t", header=None, index_col=False, usecols=[0, 1, 2],
                       names=["chrom", "start", "end"], dtype={"chrom": object})


def _make(rng, n, n_names, extra_cols=0, crlf=False, blanks=False, terminated=True, negative=False):
    names = [f"chr{i}" for i in range(1, n_names + 1)] + ["chrX", "chrUn_KI270742v1", "scaffold-12.1"]
    lines = []
    for _ in range(n):
        s = int(rng.integers(-500 if negative else 0, 3_000_000_000))
        e = s + int(rng.integers(0, 100_000))
        fields = [names[int(rng.integers(0, len(names)))], str(s), str(e)]
        fields += [f"x{int(rng.integers(0, 99))}" for _ in range(extra_cols)]
        lines.append("\t".join(fields))
        if blanks and rng.random() < 0.05:
            lines.append("")
    sep = "\r\n" if crlf else "\n"
    text = sep.join(lines) + (sep if terminated else "")
    return text.encode()


@pytest.mark.parametrize("kw", [
    dict(n=1, n_names=1), dict(n=7, n_names=2, terminated=False), dict(n=3000, n_names=24, extra_cols=3),
    dict(n=5000, n_names=300, crlf=True), dict(n=4000, n_names=5, blanks=True, negative=True),
    dict(n=200_000, n_names=24, extra_cols=1), dict(n=50_000, n_names=3000, terminated=False, blanks=True),
])
def test_bed_ingest_matches_pandas(kw):
    from bioframe_b200.io import read_bed_device

    rng = np.random.default_rng(kw["n"])
    text = _make(rng, **kw)
    got = read_bed_device(text)
    want = _reference_frame(text)
    assert len(got) == len(want)
    pd.testing.assert_frame_equal(got.to_frame(), want)
    # codes are ranks in sorted name order, ready for the engine
    codes, keys = pd.factorize(want["chrom"], sort=True)
    assert got.names == list(keys)
    assert np.array_equal(got.chrom.cpu().numpy(), codes.astype(np.int32))


def test_bed_ingest_feeds_the_engine():
    from bioframe_b200 import engine
    from bioframe_b200.io import read_bed_device
    from oracle import ops_oracle as oo

    rng = np.random.default_rng(4)
    a = read_bed_device(_make(rng, 20_000, 5))
    b = read_bed_device(_make(rng, 3_000, 5))
    assert a.names == b.names
    ev1, ev2, _ = engine.overlap_intidxs(a.chrom, a.start, a.end, b.chrom, b.start, b.end, len(a.names), how="left")
    want = oo.overlap_intidxs_coded(a.chrom.cpu().numpy(), a.start.cpu().numpy(), a.end.cpu().numpy(),
                                    b.chrom.cpu().numpy(), b.start.cpu().numpy(), b.end.cpu().numpy(), len(a.names), how="left")
    assert np.array_equal(ev1.cpu().numpy(), want[0]) and np.array_equal(ev2.cpu().numpy(), want[1])


@pytest.mark.parametrize("bad", [b"chr1\t10\n", b"chr1\t10\tabc\n", b"\t1\t2\n", b"chr1\t1\t2\nchr2 5 9\n",
                                 b"chr1\t99999999999999999999\t5\n"])
def test_bed_ingest_rejects_malformed_lines(bad):
    from bioframe_b200.io import read_bed_device

    with pytest.raises(ValueError):
        read_bed_device(bad)


def test_bed_ingest_empty_and_blank_only():
    from bioframe_b200.io import read_bed_device

    assert len(read_bed_device(b"")) == 0
    assert len(read_bed_device(b"\n\r\n\n")) == 0
