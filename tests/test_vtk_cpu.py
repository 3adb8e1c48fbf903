"""VTK output (reference: src/visualization/write2vtk.jl:113-345,
src/callbacks_step/save_solution_vtk.jl:58-173): file naming, array names / component counts, the .pvd collection, the
parallel index file and the save schedule -- all host code, checked on the CPU with a stand-in semidiscretization."""
import os
import types
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import cases


def _mft():
    import mft_b200

    return mft_b200


class _Cache:
    def __init__(self, n):
        rng = np.random.default_rng(5)
        self.eps, self.eps_uw, self.eps_rv = rng.random(n), rng.random(n), rng.random(n)
        self.eps_c = rng.integers(0, 3, n).astype(np.float64)
        self.approx_du, self.residual = rng.random((4, n)), rng.random((4, n))


def _semi(n=57, with_source=True):
    m = _mft()
    rng = np.random.default_rng(1)
    pts = rng.random((n, 2)) * 3.0
    src = None
    if with_source:
        src = object.__new__(m.SourceResidualViscosityTominec)
        src.__dict__["_cache"] = _Cache(n)
        type(src).cache = property(lambda self: self.__dict__["_cache"])
    st = types.SimpleNamespace(values=lambda: [src] if src is not None else [])
    return types.SimpleNamespace(equations=m.CompressibleEulerEquations2D(1.4), source_terms=st,
                                 domain=types.SimpleNamespace(pd=types.SimpleNamespace(points=pts))), pts


def _state(n):
    rng = np.random.default_rng(2)
    rho = 1.0 + rng.random(n)
    return np.stack([rho, rho * 0.3, -rho * 0.2, 2.5 + rng.random(n)])


def test_trixi2vtk_fields_names_and_round_trip(tmp_path):
    m = _mft()
    semi, pts = _semi()
    n = pts.shape[0]
    u = _state(n)
    f = m.trixi2vtk(u, semi, 0.25, iter=3, output_directory=str(tmp_path), prefix="run", kinetic=lambda v, uu, t, s: 0.5 * uu[1] ** 2 / uu[0])
    assert os.path.basename(f) == "run_CompressibleEulerEquations2D_1_3.vtu"       # write2vtk.jl:151-156, system_names :52-57
    points, pdata, fdata, cells = m.vtk.read_vtu(f)
    assert np.array_equal(points[:, :2], pts) and np.all(points[:, 2] == 0.0)
    assert np.array_equal(cells["types"], np.ones(n, dtype=np.uint8))              # VTK_VERTEX
    assert np.array_equal(cells["connectivity"], np.arange(n)) and np.array_equal(cells["offsets"], np.arange(1, n + 1))
    # write2vtk!(::CompressibleEulerEquations2D) :310-326
    assert np.array_equal(pdata["density"], u[0]) and np.array_equal(pdata["density_energy"], u[3])
    assert pdata["momentum"].shape == (n, 2) and np.array_equal(pdata["momentum"], u[1:3].T)
    v = u[1:3] / u[0]
    assert np.array_equal(pdata["velocity"], v.T)
    np.testing.assert_allclose(pdata["pressure"], 0.4 * (u[3] - 0.5 * (u[1] * v[0] + u[2] * v[1])), rtol=1e-15)
    # write2vtk!(::SourceResidualViscosityTominec) :335-345
    c = list(semi.source_terms.values())[0].cache
    for name, ref in (("eps", c.eps), ("eps_scalar", c.eps_c), ("eps_uw", c.eps_uw), ("eps_rv", c.eps_rv)):
        assert np.array_equal(pdata[name], ref)
    assert np.array_equal(pdata["approx_du"], c.approx_du.T) and np.array_equal(pdata["residual"], c.residual.T)
    assert np.array_equal(pdata["index"], np.arange(1, n + 1))
    np.testing.assert_allclose(pdata["kinetic"], 0.5 * u[1] ** 2 / u[0])
    assert fdata["time"][0] == 0.25 and fdata["solver_version"] == "mft_b200"
    # collection
    pvd = ET.parse(os.path.join(tmp_path, "run_CompressibleEulerEquations2D_1.pvd")).getroot()
    ds = list(pvd.iter("DataSet"))
    assert len(ds) == 1 and ds[0].get("file") == os.path.basename(f) and float(ds[0].get("timestep")) == 0.25
    # appended to when iter > 0, reset at iter 0 (paraview_collection(...; append = iter > 0))
    m.trixi2vtk(u, semi, 0.5, iter=4, output_directory=str(tmp_path), prefix="run")
    assert len(list(ET.parse(os.path.join(tmp_path, "run_CompressibleEulerEquations2D_1.pvd")).getroot().iter("DataSet"))) == 2
    m.trixi2vtk(u, semi, 0.0, iter=0, output_directory=str(tmp_path), prefix="run")
    assert len(list(ET.parse(os.path.join(tmp_path, "run_CompressibleEulerEquations2D_1.pvd")).getroot().iter("DataSet"))) == 1


def test_max_coordinates_clips_and_parallel_pieces(tmp_path, capsys):
    m = _mft()
    semi, pts = _semi(with_source=False)
    n = pts.shape[0]
    u = _state(n)
    f = m.trixi2vtk(u, semi, 0.0, iter=0, output_directory=str(tmp_path), max_coordinates=1.5)
    assert "clipped" in capsys.readouterr().out
    points, _, _, _ = m.vtk.read_vtu(f)
    assert points.max() <= 1.5
    # two ranks: owned points only, rank field, .pvtu index written by rank 0
    for r in (1, 0):
        out = m.trixi2vtk(u, semi, 1.0, iter=2, output_directory=str(tmp_path / "par"), rank=r, nranks=2, n_owned=n - 7)
    assert out.endswith("CompressibleEulerEquations2D_1_2.pvtu")
    root = ET.parse(out).getroot()
    assert [p.get("Source") for p in root.iter("Piece")] == ["CompressibleEulerEquations2D_1_2_1.vtu", "CompressibleEulerEquations2D_1_2_2.vtu"]
    names = {e.get("Name") for e in root.find(".//PPointData")}
    assert {"density", "momentum", "pressure", "velocity", "index", "rank"} <= names
    _, pdata, _, _ = m.vtk.read_vtu(str(tmp_path / "par" / "CompressibleEulerEquations2D_1_2_2.vtu"))
    assert pdata["rank"].shape == (n - 7,) and np.all(pdata["rank"] == 1)


def test_solution_saving_callback_schedule(tmp_path):
    m = _mft()
    with pytest.raises(ValueError):
        m.SolutionSavingCallback(interval=2, dt=0.1)
    cb = m.SolutionSavingCallback(interval=3, output_directory=str(tmp_path))
    due = [it for it in range(0, 11) if cb.due(0.1 * it, it, it == 10)]
    assert due == [0, 3, 6, 9, 10]                      # initial, every 3rd step, final
    cb = m.SolutionSavingCallback(dt=0.25, save_initial_solution=False, save_final_solution=False)
    assert [it for it in range(0, 11) if cb.due(0.1 * it, it, it == 10)] == [3, 5, 8, 10]
    cb = m.SolutionSavingCallback(save_times=[0.15, 0.6], save_final_solution=False)
    assert [it for it in range(0, 11) if cb.due(0.1 * it, it, it == 10)] == [0, 2, 6]
    semi, pts = _semi(with_source=False)
    cb = m.SolutionSavingCallback(interval=2, output_directory=str(tmp_path), prefix="s")
    u = _state(pts.shape[0])
    for it in (0, 2, 4):
        cb(u, semi, 0.1 * it, it, finished=False)
    cb(u, semi, 0.4, 4, finished=True)                  # final save at an already saved iteration -> iter + 1 (:151-156)
    assert [os.path.basename(f) for f in cb.files] == [f"s_CompressibleEulerEquations2D_1_{i}.vtu" for i in (0, 2, 4, 5)]
