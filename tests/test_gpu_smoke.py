"""`__graft_entry__.smoke()` -- the driver's one small invocation of the hot path on cuda:0, checked against the
oracle -- kept green by the suite itself."""
import pytest


@pytest.mark.gpu
def test_smoke_entry_point(capsys):
    import __graft_entry__ as entry

    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
