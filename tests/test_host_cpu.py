"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol of include/rife_b200.h, argument
validation works without a device, and the model loader matches the file-format arithmetic."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "rife_b200.h")).read()
    declared = set(re.findall(r"\b(rife_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg.EXPORTS), declared ^ set(pkg.EXPORTS)
    L = pkg.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_create_rejects_cpu_mode_and_bad_args(pkg):
    L = pkg.lib()
    h = ctypes.c_void_p()
    assert L.rife_b200_create(ctypes.byref(h), -1, 0, 0, 0, 1, 0, 1) == -1  # the reference's -g -1: no CPU fallback
    assert not h.value
    assert L.rife_b200_create(None, 0, 0, 0, 0, 1, 0, 1) == -1
    with pytest.raises(pkg.RifeError):
        pkg.RIFE(-1)


def test_family_flags_match_reference_sniffing(pkg):
    assert pkg.family_flags("models/rife-v4.6") == (False, True)
    assert pkg.family_flags("rife-v2.3") == (True, False)
    assert pkg.family_flags("rife-v3.1") == (True, False)
    assert pkg.family_flags("rife-anime") == (False, False)
    assert pkg.family_flags("rife-HD/") == (False, False)


def test_synth_stream_equals_frame_by_frame():
    """bench.py builds its frame stream from one canvas (synth.stream); it must be the frames synth.frame defines."""
    import numpy as np
    import parity
    for k0, n, w, h, dx, dy, seed in ((0, 9, 64, 48, 3, 2, 0), (5, 7, 70, 50, 3, 2, 1), (2, 5, 64, 48, -3, 2, 0), (0, 4, 33, 31, 24, 16, 2), (3, 3, 40, 40, 0, 0, 0)):
        fr = parity.synth.stream(k0, n, w, h, dx, dy, seed)
        assert len(fr) == n
        for i in range(n):
            assert np.array_equal(fr[i], parity.synth.frame(k0 + i, w, h, dx, dy, seed)), (k0, i)
            assert fr[i].flags["C_CONTIGUOUS"]
