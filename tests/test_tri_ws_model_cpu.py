"""The producer / consumer protocol of the warp-specialised tridiagonal sweep kernel (csrc/tridiag.cu: k_tri_sweep_ws)
under random interleavings of its three actors (tools/tri_ws_model.py): no deadlock, boxes consumed once and in order,
no stage overwritten before it is released or read while a load is in flight."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import tri_ws_model  # noqa: E402


@pytest.mark.parametrize("ns", [2, 8, 16, 32])
def test_ring_protocol_under_random_interleavings(ns):
    for nb in (1, 2, 7, 8, 9, 33, 129):
        for seed in range(10):
            tri_ws_model.simulate(nb, ns, seed)


def test_a_ring_of_one_is_rejected():
    with pytest.raises(AssertionError):
        tri_ws_model.simulate(4, 1, 0)
