"""Drop-in for ``pyaxisymflow.utils.dump_vtk`` (``utils/dump_vtk.py:5-62``) without the ``vtk`` package.

The reference fills one multi-component ``vtkDoubleArray`` called ``"softy"`` (component i = field i, named
after ``field_names_list[i]``) on a ``vtkImageData`` of dimensions ``(grid_size_z, grid_size_r, 1)`` and writes it
with ``vtkXMLImageDataWriter``.  ``vtk_write`` here produces the same data set as a VTK XML ImageData (``.vti``)
file -- same extent, origin, spacing, array name, component names and values; raw appended encoding instead of
the library's zlib blocks -- so ParaView / ``vtkXMLImageDataReader`` open it the same way.  Fields may be NumPy
arrays, CUDA tensors or ``DeviceField``s; device fields are fetched through ``io.FieldSnapshotter`` (pinned
buffers, side stream), and with ``writer.asynchronous = True`` the file is written by a worker thread while the
loop goes on.  ``read_vti`` is the matching reader used by the tests.
"""
from __future__ import annotations

import re
import struct

import numpy as np


class VtiWriter:
    """stands in for the (vtk_image_data, temp_vtk_array, writer) triple of the reference"""

    def __init__(self, grid_size_z, grid_size_r):
        self.grid_size_z, self.grid_size_r = int(grid_size_z), int(grid_size_r)
        self.asynchronous = False
        self._snap = None

    def snapshotter(self):
        if self._snap is None:
            from ..io import FieldSnapshotter

            self._snap = FieldSnapshotter()
        return self._snap

    def wait(self):
        if self._snap is not None:
            self._snap.wait()


def vtk_init(grid_size_z, grid_size_r):
    """utils/dump_vtk.py:5-25 -- returns the same 3-tuple shape the drivers unpack."""
    w = VtiWriter(grid_size_z, grid_size_r)
    return w, None, w


def write_vti(filename, names, arrays, grid_size_z, grid_size_r, array_name="softy"):
    """one Float64 array with len(names) components over a (grid_size_z, grid_size_r, 1) image"""
    n = grid_size_z * grid_size_r
    comps = []
    for a in arrays:
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if a.size != n:
            raise ValueError(f"field has {a.size} values, image has {n} points")
        comps.append(a)
    data = np.stack(comps, axis=1).reshape(-1) if comps else np.zeros(0)      # tuple-major, like CopyComponent
    attrs = " ".join(f'ComponentName{i}="{nm}"' for i, nm in enumerate(names))
    ext = f"0 {grid_size_z - 1} 0 {grid_size_r - 1} 0 0"
    header = (
        '<?xml version="1.0"?>\n'
        '<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
        f'  <ImageData WholeExtent="{ext}" Origin="0 0 0" Spacing="1 1 1">\n'
        f'    <Piece Extent="{ext}">\n'
        '      <PointData>\n'
        f'        <DataArray type="Float64" Name="{array_name}" NumberOfComponents="{len(names)}" {attrs} '
        'format="appended" offset="0"/>\n'
        '      </PointData>\n'
        '      <CellData/>\n'
        '    </Piece>\n'
        '  </ImageData>\n'
        '  <AppendedData encoding="raw">\n   _'
    )
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(struct.pack("<Q", data.nbytes))
        f.write(data.astype("<f8", copy=False).tobytes())
        f.write(b"\n  </AppendedData>\n</VTKFile>\n")


def read_vti(filename):
    """-> (dims (nz, nr), array name, {component name: (nr, nz) array}) of a file written by ``write_vti``"""
    raw = open(filename, "rb").read()
    cut = raw.index(b'<AppendedData encoding="raw">')
    head = raw[:cut].decode("ascii")
    ext = [int(x) for x in re.search(r'WholeExtent="([^"]+)"', head).group(1).split()]
    nz, nr = ext[1] + 1, ext[3] + 1
    name = re.search(r'<DataArray[^>]*Name="([^"]+)"', head).group(1)
    ncomp = int(re.search(r'NumberOfComponents="(\d+)"', head).group(1))
    names = [re.search(rf'ComponentName{i}="([^"]*)"', head).group(1) for i in range(ncomp)]
    start = raw.index(b"_", cut) + 1
    (nbytes,) = struct.unpack("<Q", raw[start:start + 8])
    data = np.frombuffer(raw[start + 8:start + 8 + nbytes], dtype="<f8").reshape(nr * nz, ncomp)
    out = {}
    for i, nm in enumerate(names):
        out.setdefault(nm, data[:, i].reshape(nr, nz).copy())
    return (nz, nr), name, out


def vtk_write(filename, vtk_image_data, temp_vtk_array, writer, field_names_list, field_list, grid_size_z,
              grid_size_r):
    """utils/dump_vtk.py:28-62, same arguments."""
    w = writer if isinstance(writer, VtiWriter) else vtk_image_data
    if not isinstance(w, VtiWriter):
        raise TypeError("vtk_write needs the objects returned by pyaxisymflow_b200.utils.dump_vtk.vtk_init")
    if len(field_names_list) != len(field_list):
        raise ValueError("field_names_list and field_list differ in length")
    names = list(field_names_list)

    def sink(host):
        write_vti(filename, names, [host[f"f{i}"] for i in range(len(names))], grid_size_z, grid_size_r)

    on_device = any(not isinstance(f, np.ndarray) for f in field_list)
    if on_device or w.asynchronous:
        snap = w.snapshotter()
        snap.snapshot({f"f{i}": f for i, f in enumerate(field_list)}, sink)
        if not w.asynchronous:
            snap.wait()
    else:
        sink({f"f{i}": f for i, f in enumerate(field_list)})
