"""__graft_entry__.smoke() — the call the driver makes on the GPU box — run here against the kernels on the SIMT emulator, so that
the smoke test itself (inputs, expectations, the binding it drives) is known to be sound before it meets a GPU."""
import os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_smoke_logic_on_the_emulator(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import genozip_b200
    from simt_lib import simt_engine_class
    monkeypatch.setattr(genozip_b200, "Engine", simt_engine_class())
    import __graft_entry__ as g
    g.smoke()
    assert "smoke ok" in capsys.readouterr().out
