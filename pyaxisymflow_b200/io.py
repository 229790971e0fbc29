"""Output on either side of the timestep loop: restart files and field dumps taken from
device-resident fields without stalling the loop.

The reference's drivers write ``np.savez("restart.npz", t=..., vorticity=..., ...)`` and reload it with
``np.load`` (``examples/ParticleOscillatoryFlowCases/particle_in_bubble_oscillatory_flow.py:129-147, 236-257``),
and dump ``.vti`` images through ``utils/dump_vtk.py``.  Here a :class:`FieldSnapshotter` copies the fields into
device staging buffers on the compute stream (an HBM-speed copy, ordered with the loop's kernels, so it sees a
consistent step and the loop may overwrite the fields at once), moves those to pinned host buffers on a side
stream, and hands them to a worker thread that does the file I/O; the compute stream is never synchronised.  ``save_npz`` / ``load_npz`` keep the reference's on-disk format: a plain ``.npz`` with the same
keys, readable by ``np.load`` on either side.  ``save_restart`` / ``load_restart`` of the device-resident steppers
use a PRIVATE key set (``_STATE`` below: the stepper's own fields and loop scalars -- e.g. ``part_char_func`` rather
than the reference's ``part_phi``, no trajectory lists): they round-trip a stepper bit for bit but are not
interchangeable with a ``restart.npz`` written by the reference's driver; ``load_restart`` names the missing keys.
"""
from __future__ import annotations

import queue
import threading

import numpy as np
import torch

from .device import DeviceField


def _as_tensor(x):
    if isinstance(x, DeviceField):
        return x.t
    return x


class FieldSnapshotter:
    """Asynchronous device->host snapshots.  ``snapshot(items, sink)`` returns at once; ``sink(host_dict)`` runs in
    the worker thread when the copies have landed.  ``depth`` snapshots may be in flight (pinned buffers are
    recycled per (shape, dtype)); a further call blocks until one has been written."""

    def __init__(self, depth=2):
        self.depth = int(depth)
        self._cuda = torch.cuda.is_available()
        self._sides = {}                      # one side stream per device
        self._pool = {}
        self._slots = threading.Semaphore(self.depth)
        self._q = queue.Queue()
        self._err = None
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    def _buffer(self, t):
        key = (tuple(t.shape), t.dtype, "host")
        free = self._pool.setdefault(key, [])
        if free:
            return free.pop()
        return torch.empty(t.shape, dtype=t.dtype, pin_memory=self._cuda)

    def _dev_buffer(self, t):
        key = (tuple(t.shape), t.dtype, "dev", t.device)
        free = self._pool.setdefault(key, [])
        if free:
            return free.pop()
        return torch.empty(t.shape, dtype=t.dtype, device=t.device)

    def snapshot(self, items, sink):
        if self._err is not None:
            err, self._err = self._err, None
            raise err
        self._slots.acquire()
        try:
            job = self._stage(items, sink)
        except BaseException:
            self._slots.release()
            raise
        self._q.put(job)

    def _stage(self, items, sink):
        host, used, used_dev = {}, [], []
        event = None
        dev = {k: _as_tensor(v) for k, v in items.items()}
        cuda_items = {k: v for k, v in dev.items() if isinstance(v, torch.Tensor) and v.is_cuda}
        if cuda_items:
            # 1) device-to-device copy into a staging buffer ON THE COMPUTE STREAM (HBM speed, ordered with the
            #    loop's kernels: later steps may overwrite the field at once), 2) the slow device-to-host copy of
            #    the staging buffer on the side stream, overlapped with whatever the loop does next
            stage = {}
            for k, v in cuda_items.items():
                st = self._dev_buffer(v)
                st.copy_(v)
                stage[k] = st
            device = next(iter(cuda_items.values())).device
            side = self._sides.get(device)
            if side is None:
                side = self._sides[device] = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for k, st in stage.items():
                    buf = self._buffer(st)
                    buf.copy_(st, non_blocking=True)
                    used.append(buf)
                    used_dev.append(st)
                    host[k] = buf
                event = torch.cuda.Event()
                event.record(side)
        for k, v in dev.items():
            if k in host:
                continue
            if isinstance(v, torch.Tensor):
                host[k] = v.clone()
            elif isinstance(v, np.ndarray):
                host[k] = v.copy()           # the caller may overwrite it before the worker runs
            else:
                host[k] = v
        return host, used + used_dev, event, sink

    def _run(self):
        while True:
            job = self._q.get()
            if job is None:
                return
            host, used, event, sink = job
            try:
                if event is not None:
                    event.synchronize()
                sink({k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in host.items()})
            except BaseException as e:       # surfaced by the next snapshot() / wait()
                self._err = e
            finally:
                for buf in used:
                    key = ((tuple(buf.shape), buf.dtype, "dev", buf.device) if buf.is_cuda
                           else (tuple(buf.shape), buf.dtype, "host"))
                    self._pool.setdefault(key, []).append(buf)
                self._slots.release()
                self._q.task_done()

    def wait(self):
        self._q.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def close(self):
        self.wait()
        self._q.put(None)


_default = None


def _snapshotter():
    global _default
    if _default is None:
        _default = FieldSnapshotter()
    return _default


def save_npz(path, asynchronous=False, **items):
    """``np.savez(path, **items)`` where items may live on the GPU.  With ``asynchronous`` the call returns once the
    copies are queued; ``wait()`` (or the next save) surfaces I/O errors."""
    snap = _snapshotter()
    snap.snapshot(items, lambda host: np.savez(path, **host))
    if not asynchronous:
        snap.wait()


def wait():
    if _default is not None:
        _default.wait()


def load_npz(path):
    """dict of NumPy arrays / scalars, like ``np.load`` (0-d arrays stay 0-d, as the drivers ``float(...)`` them)"""
    with np.load(path) as f:
        return {k: f[k] for k in f.files}


# --------------------------------------------------------------------------------------
# restart files of the device-resident steppers (keys follow the reference's restart.npz where it has them)
# --------------------------------------------------------------------------------------
_STATE = {
    "RigidFlowStepper": (["vorticity", "state"], []),           # state = the 8 device-resident loop scalars
    "SoftSphereStepper": (["vorticity", "eta1", "eta2", "ball_phi", "avg_psi", "avg_phi"], ["t", "freqTimer", "it"]),
    "ParticleFlowStepper": (["vorticity", "part_char_func", "avg_psi", "avg_vort", "avg_part_char_func"],
                            ["t", "it", "U_z_cm_part", "diff", "part_Z_cm", "F_total", "freqTimer", "avg_Z_cm",
                             "avg_time"]),
}


def save_restart(stepper, path="restart.npz", asynchronous=False):
    """particle_in_bubble_oscillatory_flow.py:236-257 for a device-resident stepper"""
    fields, scalars = _STATE[type(stepper).__name__]
    if getattr(stepper, "device_scalars", False):
        stepper.sync_scalars()                                   # device-resident loop scalars -> host attributes
    items = {k: getattr(stepper, k) for k in fields}
    if type(stepper).__name__ == "RigidFlowStepper":
        items["t"] = stepper.state[0:1]                         # the key the reference's restart files lead with
    for k in scalars:
        items[k] = getattr(stepper, k)
    save_npz(path, asynchronous=asynchronous, **items)


def load_restart(stepper, path="restart.npz"):
    """particle_in_bubble_oscillatory_flow.py:129-147: fields are copied in place, scalars restored"""
    data = load_npz(path)
    fields, scalars = _STATE[type(stepper).__name__]
    missing = [k for k in fields + scalars if k not in data]
    if missing:
        raise KeyError(f"{path} is not a restart file of {type(stepper).__name__}: missing {missing} "
                       "(stepper restarts use their own key set, see the module docstring)")
    for k in fields:
        getattr(stepper, k).copy_(torch.from_numpy(np.ascontiguousarray(data[k])))
    for k in scalars:
        cur = getattr(stepper, k)
        setattr(stepper, k, type(cur)(data[k]))
    if getattr(stepper, "device_scalars", False):
        stepper.push_scalars()                                   # ... and back into the device block
    return stepper
