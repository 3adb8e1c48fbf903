"""The usage example (examples/cylinder_euler_rv.py: the reference's driver script with the CUDA engine) runs end to end:
adaptive SSPRK43, history callback, residual viscosity, VTK snapshots.  (First hardware run: round-end pass.)"""
import os
import runpy
import sys

import pytest

import cases

pytestmark = pytest.mark.gpu


def test_cylinder_example_runs(tmp_path, capsys, monkeypatch):
    out = str(tmp_path / "out")
    monkeypatch.setattr(sys, "argv", ["cylinder_euler_rv.py", "--tend", "0.002", "--out", out])
    runpy.run_path(os.path.join(cases.ROOT, "examples", "cylinder_euler_rv.py"), run_name="__main__")
    text = capsys.readouterr().out
    assert "non-finite entries: 0" in text and "accepted" in text
    files = sorted(os.listdir(out))
    assert any(f.endswith(".pvd") for f in files) and sum(f.endswith(".vtu") for f in files) >= 2
