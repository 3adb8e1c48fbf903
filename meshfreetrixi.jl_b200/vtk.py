"""VTK output of the resident state: the on-disk format the reference's callers consume.

  trixi2vtk               src/visualization/write2vtk.jl:113-250  (one .vtu of VTK_VERTEX cells per save + a .pvd collection;
                          .pvtu + one piece per rank under MPI)
  write2vtk! field sets   src/visualization/write2vtk.jl:306-342  (density, density_energy, momentum, pressure, velocity;
                          eps, eps_scalar, eps_uw, eps_rv, approx_du, residual for the viscosity sources)
  SolutionSavingCallback  src/callbacks_step/save_solution_vtk.jl:58-173

The reference delegates the byte layout to WriteVTK.jl (third party); here the files are written directly as XML
UnstructuredGrid with raw appended data (UInt64 block headers, little endian), which ParaView / VTK / meshio read.  Array
names, component counts, file names (`<prefix_>CompressibleEulerEquations2D_1[_iter].vtu`, collection `….pvd`) and the
field data (`time`, `solver_version`) follow the reference.  Host code: the state comes from `mft_download_state` /
`mft_get_field`; nothing here is on the timed path.
"""
from __future__ import annotations

import os
import re
import struct
import xml.etree.ElementTree as ET

import numpy as np

_VTK_TYPES = {np.dtype("float64"): "Float64", np.dtype("float32"): "Float32", np.dtype("int64"): "Int64",
              np.dtype("int32"): "Int32", np.dtype("uint8"): "UInt8"}
_NP_TYPES = {v: k for k, v in _VTK_TYPES.items()}


def _as_point_array(a, n):
    """(n,), (c, n) [the reference's component-major layout] or (n, c) -> (n, c) C-contiguous"""
    a = np.asarray(a)
    if a.ndim == 1:
        assert a.shape[0] == n
        return np.ascontiguousarray(a.reshape(n, 1))
    if a.shape[0] != n and a.shape[1] == n:
        a = a.T
    assert a.shape[0] == n, f"array of shape {a.shape} does not match {n} points"
    return np.ascontiguousarray(a)


def write_vtu(path, points, point_data=None, field_data=None):
    """One UnstructuredGrid piece: every point is a VTK_VERTEX cell (write2vtk.jl:163-164).  points: (N, 2|3)."""
    points = np.asarray(points, dtype=np.float64)
    n = points.shape[0]
    p3 = np.zeros((n, 3))
    p3[:, :points.shape[1]] = points
    blocks, offset = [], 0

    def add(arr):
        nonlocal offset
        raw = np.ascontiguousarray(arr).tobytes()
        off = offset
        blocks.append(struct.pack("<Q", len(raw)) + raw)
        offset += 8 + len(raw)
        return off

    def darray(name, arr, ncomp, extra=""):
        return (f'<DataArray type="{_VTK_TYPES[arr.dtype]}" Name="{name}" NumberOfComponents="{ncomp}" format="appended" '
                f'offset="{add(arr)}"{extra}/>')

    out = ['<?xml version="1.0"?>',
           '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
           "<UnstructuredGrid>"]
    if field_data:
        out.append("<FieldData>")
        for k, v in field_data.items():
            if isinstance(v, str):
                b = np.frombuffer(v.encode() + b"\0", dtype=np.uint8)
                out.append(f'<DataArray type="UInt8" Name="{k}" NumberOfTuples="1" NumberOfComponents="{len(b)}" '
                           f'format="appended" offset="{add(b)}"/>')   # WriteVTK stores strings as one tuple of chars
            else:
                a = np.atleast_1d(np.asarray(v, dtype=np.float64))
                out.append(darray(k, a, 1, f' NumberOfTuples="{len(a)}"'))
        out.append("</FieldData>")
    out.append(f'<Piece NumberOfPoints="{n}" NumberOfCells="{n}">')
    out.append("<Points>" + darray("Points", p3, 3) + "</Points>")
    out.append("<Cells>")
    out.append(darray("connectivity", np.arange(n, dtype=np.int64), 1))
    out.append(darray("offsets", np.arange(1, n + 1, dtype=np.int64), 1))
    out.append(darray("types", np.ones(n, dtype=np.uint8), 1))          # VTK_VERTEX = 1
    out.append("</Cells>")
    out.append("<PointData>")
    for k, v in (point_data or {}).items():
        a = _as_point_array(v, n)
        if a.dtype not in _VTK_TYPES:
            a = a.astype(np.float64)
        out.append(darray(k, a, a.shape[1]))
    out.append("</PointData>")
    out.append("</Piece></UnstructuredGrid>")
    out.append('<AppendedData encoding="raw">')
    head = ("\n".join(out) + "\n_").encode()
    with open(path, "wb") as f:
        f.write(head)
        for b in blocks:
            f.write(b)
        f.write(b"\n</AppendedData>\n</VTKFile>\n")
    return path


def read_vtu(path):
    """Reader for the files write_vtu produces (tests, post-processing): returns points (N,3), point_data, field_data."""
    raw = open(path, "rb").read()
    m = re.search(rb'<AppendedData encoding="raw">\s*_', raw)
    xml_part = raw[:m.start()] + b"</VTKFile>"
    data = raw[m.end():]
    root = ET.fromstring(xml_part)

    def load(el):
        off = int(el.get("offset"))
        nbytes = struct.unpack_from("<Q", data, off)[0]
        a = np.frombuffer(data, dtype=_NP_TYPES[el.get("type")], count=nbytes // _NP_TYPES[el.get("type")].itemsize,
                          offset=off + 8)
        nc = int(el.get("NumberOfComponents", "1"))
        return a.reshape(-1, nc) if nc > 1 else a

    piece = root.find(".//Piece")
    points = load(piece.find("Points/DataArray"))
    pdata = {el.get("Name"): load(el) for el in piece.findall("PointData/DataArray")}
    fdata = {}
    for el in root.findall(".//FieldData/DataArray"):
        a = load(el)
        fdata[el.get("Name")] = bytes(a.reshape(-1)).rstrip(b"\0").decode() if el.get("type") == "UInt8" else a
    cells = {el.get("Name"): load(el) for el in piece.findall("Cells/DataArray")}
    return points, pdata, fdata, cells


def write_pvtu(path, piece_files, point_arrays, field_arrays=()):
    """Parallel index file (pvtk_grid, write2vtk.jl:188-219): point_arrays = [(name, vtk type, ncomp)]"""
    out = ['<?xml version="1.0"?>', '<VTKFile type="PUnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
           '<PUnstructuredGrid GhostLevel="0">']
    if field_arrays:
        out.append("<PFieldData>" + "".join(f'<PDataArray type="{t}" Name="{k}" NumberOfComponents="{c}"/>' for k, t, c in field_arrays)
                   + "</PFieldData>")
    out.append('<PPoints><PDataArray type="Float64" Name="Points" NumberOfComponents="3"/></PPoints>')
    out.append("<PPointData>" + "".join(f'<PDataArray type="{t}" Name="{k}" NumberOfComponents="{c}"/>' for k, t, c in point_arrays)
               + "</PPointData>")
    for pf in piece_files:
        out.append(f'<Piece Source="{os.path.basename(pf)}"/>')
    out.append("</PUnstructuredGrid></VTKFile>")
    open(path, "w").write("\n".join(out) + "\n")
    return path


class PvdCollection:
    """paraview_collection(file; append = iter > 0)  (write2vtk.jl:159): time -> file index"""

    def __init__(self, path, append=True):
        self.path = path if path.endswith(".pvd") else path + ".pvd"
        self.entries = []
        if append and os.path.exists(self.path):
            for el in ET.parse(self.path).getroot().iter("DataSet"):
                self.entries.append((float(el.get("timestep")), el.get("file")))

    def add(self, t, file):
        self.entries.append((float(t), os.path.basename(file)))

    def save(self):
        out = ['<?xml version="1.0"?>', '<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">', "<Collection>"]
        out += [f'<DataSet timestep="{t!r}" part="0" file="{f}"/>' for t, f in self.entries]
        out += ["</Collection>", "</VTKFile>"]
        open(self.path, "w").write("\n".join(out) + "\n")
        return self.path


def euler_fields(u, gamma):
    """write2vtk!(vtk, u, t, ::CompressibleEulerEquations2D, ...) (write2vtk.jl:310-326): conservative + primitive fields;
    cons2prim as Trixi computes it (v = m / rho, p = (gamma-1)(E - (m1 v1 + m2 v2)/2))"""
    rho, m1, m2, E = u
    v1, v2 = m1 / rho, m2 / rho
    p = (gamma - 1.0) * (E - 0.5 * (m1 * v1 + m2 * v2))
    return {"density": rho, "density_energy": E, "momentum": np.stack([m1, m2]), "pressure": p, "velocity": np.stack([v1, v2])}


def source_fields(source):
    """write2vtk! for SourceUpwindViscosityTominec / SourceResidualViscosityTominec (write2vtk.jl:328-345); the fields live
    on the device and are fetched with mft_get_field (needs the engine's diagnostics option)"""
    name = type(source).__name__
    if name == "SourceIGR":          # write2vtk.jl:344-349
        return {"sigma": source.cache.sigma}
    if name not in ("SourceUpwindViscosityTominec", "SourceResidualViscosityTominec"):
        return {}
    c = source.cache
    out = {"eps": c.eps, "eps_scalar": c.eps_c}
    if name == "SourceResidualViscosityTominec":
        out.update({"eps_uw": c.eps_uw, "eps_rv": c.eps_rv, "approx_du": c.approx_du, "residual": c.residual})
    return out


def trixi2vtk(u, semi, t, iter=None, output_directory="out", prefix="", write_meta_data=True, max_coordinates=np.inf,
              rank=None, nranks=1, n_owned=None, **custom_quantities):
    """trixi2vtk(u_ode, semi, t; iter, output_directory, prefix, write_meta_data, max_coordinates, custom_quantities...)
    (write2vtk.jl:113-250).  u: (V, N) host array.  Multi-rank: every rank writes its piece (owned points only) and rank 0
    the .pvtu index + collection.  Returns the path of the file ParaView should open."""
    os.makedirs(output_directory, exist_ok=True)
    eq = semi.equations
    system_name = type(eq).__name__ + "_1"
    pre = "" if prefix == "" else f"{prefix}_"
    post = "" if iter is None else f"_{iter}"
    base = os.path.join(output_directory, pre + system_name + post)
    pts = np.array(semi.domain.pd.points, dtype=np.float64)
    n = pts.shape[0] if n_owned is None else int(n_owned)
    pts = pts[:n]
    if np.abs(pts).max(initial=0.0) > max_coordinates:
        print("Warning: At least one particle's absolute coordinates exceed `max_coordinates` and have been clipped")
        pts = np.clip(pts, -max_coordinates, max_coordinates)
    u = np.asarray(u)
    fields = {}
    if type(eq).__name__ == "CompressibleEulerEquations2D":
        fields.update(euler_fields(u[:, :n], eq.gamma))
    else:
        fields["scalar"] = u[0, :n]
    for src in (semi.source_terms.values() if semi.source_terms is not None else []):
        try:
            for k, v in source_fields(src).items():
                fields[k] = np.asarray(v)[..., :n]
        except Exception as exc:   # diagnostics not enabled on the engine: the state fields are still written
            fields.setdefault("_skipped", str(exc))
    skipped = fields.pop("_skipped", None)
    for k, q in custom_quantities.items():
        val = q(u, u, t, eq) if callable(q) else q
        if val is not None:
            fields[str(k)] = np.asarray(val)[..., :n]
    fields["index"] = np.arange(1, n + 1, dtype=np.int64)       # eachelement: 1-based
    fdata = {"time": float(t)}
    if write_meta_data:
        fdata["solver_version"] = "mft_b200"
        if skipped:
            fdata["skipped_fields"] = skipped
    if nranks > 1:
        fields["rank"] = np.full(n, rank, dtype=np.int64)
        piece = f"{base}_{rank + 1}.vtu"
        write_vtu(piece, pts, fields, fdata)
        out = base + ".pvtu"
        if rank == 0:
            arrays = [(k, _VTK_TYPES.get(np.asarray(v).dtype, "Float64"), _as_point_array(v, n).shape[1]) for k, v in fields.items()]
            write_pvtu(out, [f"{base}_{r + 1}.vtu" for r in range(nranks)], arrays, [("time", "Float64", 1)])
    else:
        out = write_vtu(base + ".vtu", pts, fields, fdata)
    if rank in (None, 0):
        pvd = PvdCollection(os.path.join(output_directory, pre + system_name), append=(iter or 0) > 0)
        pvd.add(t, out)
        pvd.save()
    return out


class SolutionSavingCallback:
    """SolutionSavingCallback(; interval, dt, save_times, save_initial_solution, save_final_solution, output_directory,
    prefix, write_meta_data, max_coordinates, custom_quantities...)  (save_solution_vtk.jl:58-111).  `interval` counts
    accepted steps; `dt` saves at the first step at or past every multiple of dt (the host loop does not insert tstops)."""

    def __init__(self, interval=0, dt=0.0, save_times=(), save_initial_solution=True, save_final_solution=True,
                 output_directory="out", prefix="", write_meta_data=True, verbose=False, max_coordinates=float(2 ** 15),
                 **custom_quantities):
        if (dt > 0 and interval > 0) or (len(save_times) > 0 and (dt > 0 or interval > 0)):
            raise ValueError("Setting multiple save times for the same solution callback is not possible. "
                             "Use either `dt`, `interval` or `save_times`.")
        self.interval, self.dt, self.save_times = int(interval), float(dt), sorted(float(x) for x in save_times)
        self.save_initial_solution, self.save_final_solution = save_initial_solution, save_final_solution
        self.output_directory, self.prefix, self.write_meta_data = output_directory, prefix, write_meta_data
        self.verbose, self.max_coordinates, self.custom_quantities = verbose, max_coordinates, custom_quantities
        self.latest_saved_iter = -1
        self.files = []
        self._next_time_index = 0

    def due(self, t, it, finished):
        if it == 0:
            return self.save_initial_solution
        if finished and self.save_final_solution:
            return True
        if self.interval > 0:
            return it % self.interval == 0
        if self.dt > 0:
            k = int(np.floor(t / self.dt + 1e-12))
            if k > self._next_time_index:
                self._next_time_index = k
                return True
            return False
        if self.save_times:
            hit = False
            while self._next_time_index < len(self.save_times) and t >= self.save_times[self._next_time_index] - 1e-14:
                self._next_time_index += 1
                hit = True
            return hit
        return False

    def __call__(self, u, semi, t, it, finished=False, **parallel):
        """save now (the integrators call this after the history callbacks); `u` is the downloaded state"""
        iter_ = it if self.interval > 0 or it == 0 else (self.latest_saved_iter + 1)
        if iter_ == self.latest_saved_iter:
            iter_ += 1
        self.latest_saved_iter = iter_
        if self.verbose:
            print(f"Writing solution to {self.output_directory} at t = {t}")
        f = trixi2vtk(u, semi, t, iter=iter_, output_directory=self.output_directory, prefix=self.prefix,
                      write_meta_data=self.write_meta_data, max_coordinates=self.max_coordinates, **parallel,
                      **self.custom_quantities)
        self.files.append(f)
        return f
