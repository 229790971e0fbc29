"""Output formats on either side of the loop: the .vti writer and the restart .npz, CPU side.
The reference writes .vti through the vtk package (not installed here, and not needed by this path): the file
written here carries the same data set -- checked by reading it back -- not the same bytes."""
import os

import numpy as np
import pytest

from pyaxisymflow_b200 import io
from pyaxisymflow_b200.utils.dump_vtk import read_vti, vtk_init, vtk_write


def test_vti_round_trip(tmp_path):
    nz, nr = 12, 5
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((nr, nz)), rng.standard_normal((nr, nz))
    image, tmp, writer = vtk_init(nz, nr)
    path = os.path.join(tmp_path, "snap_0001.vti")
    vtk_write(path, image, tmp, writer, ["vorticity", "u_z"], [a, b], nz, nr)
    dims, name, fields = read_vti(path)
    assert dims == (nz, nr) and name == "softy"            # the reference's array name (utils/dump_vtk.py:41)
    assert np.array_equal(fields["vorticity"], a) and np.array_equal(fields["u_z"], b)
    head = open(path, "rb").read(600).decode("ascii", "ignore")
    assert 'type="ImageData"' in head and f'WholeExtent="0 {nz - 1} 0 {nr - 1} 0 0"' in head
    assert 'NumberOfComponents="2"' in head and 'ComponentName0="vorticity"' in head
    with pytest.raises(ValueError):
        vtk_write(path, image, tmp, writer, ["a"], [a, b], nz, nr)
    with pytest.raises(ValueError):
        vtk_write(path, image, tmp, writer, ["a"], [a[:, :-1]], nz, nr)


def test_vti_asynchronous_writer_copies_before_returning(tmp_path):
    nz, nr = 16, 8
    a = np.arange(nr * nz, dtype=np.float64).reshape(nr, nz)
    image, tmp, writer = vtk_init(nz, nr)
    writer.asynchronous = True
    path = os.path.join(tmp_path, "async.vti")
    vtk_write(path, image, tmp, writer, ["f"], [a], nz, nr)
    a[...] = -1.0                                           # the loop goes on and overwrites the field
    writer.wait()
    assert np.array_equal(read_vti(path)[2]["f"], np.arange(nr * nz, dtype=np.float64).reshape(nr, nz))


def test_restart_npz_has_the_reference_format(tmp_path):
    """particle_in_bubble_oscillatory_flow.py:236-257 / :129-147: plain np.savez keys, np.load on the other side"""
    path = os.path.join(tmp_path, "restart.npz")
    w = np.random.default_rng(1).standard_normal((6, 9))
    io.save_npz(path, t=0.125, vorticity=w, T=np.array([1.0, 2.0]), diff=0)
    with np.load(path) as f:
        assert sorted(f.files) == ["T", "diff", "t", "vorticity"]
        assert float(f["t"]) == 0.125 and np.array_equal(f["vorticity"], w) and int(f["diff"]) == 0
    got = io.load_npz(path)
    assert np.array_equal(got["vorticity"], w)
    io.save_npz(path, asynchronous=True, t=1.0, vorticity=w + 1)
    io.wait()
    assert float(np.load(path)["t"]) == 1.0


def test_snapshot_errors_surface(tmp_path):
    snap = io.FieldSnapshotter()

    def bad(host):
        raise OSError("disk full")

    snap.snapshot({"a": np.zeros(3)}, bad)
    with pytest.raises(OSError):
        snap.wait()
    snap.close()
