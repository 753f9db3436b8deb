"""The model loader (csrc/model.cpp: own parser of the reference's .param/.bin format, no ncnn) on the CPU: every network
of every model directory the reference ships parses with the counts its header declares and consumes its .bin exactly;
damaged files give RIFE_B200_ERR_MODEL with a message -- never a crash or a giant allocation."""
import ctypes
import glob
import os
import shutil

import pytest

import parity

REF_MODELS = "/root/reference/models"


def _parse(pkg, param, binf):
    L = pkg.lib()
    nl, nb, nv = ctypes.c_int(), ctypes.c_int(), ctypes.c_ulonglong()
    err = ctypes.create_string_buffer(512)
    r = L.rife_b200_debug_parse_model(os.fsencode(param), os.fsencode(binf), ctypes.byref(nl), ctypes.byref(nb), ctypes.byref(nv), err, 512)
    return r, nl.value, nb.value, nv.value, err.value.decode(errors="replace")


def _model_dirs():
    root = REF_MODELS if os.path.isdir(REF_MODELS) else os.path.join(parity.REF_DIR, "models")
    return sorted(d for d in glob.glob(os.path.join(root, "*")) if os.path.isdir(d))


def test_every_shipped_network_parses_with_its_declared_counts(pkg):
    dirs = _model_dirs()
    if not dirs:
        pytest.skip("no model directories on this host")
    checked = 0
    for d in dirs:
        for param in sorted(glob.glob(os.path.join(d, "*.param"))):
            binf = param[:-6] + ".bin"
            r, nl, nb, nv, err = _parse(pkg, param, binf)
            assert r == 0, (param, err)
            head = open(param).read().split()
            assert int(head[0]) == 7767517 and nl == int(head[1]) and nb <= int(head[2]), (param, nl, nb, head[:3])
            # fp16 weights + fp32 biases / slopes + one 4-byte tag per weighted layer: the value count bounds the file size
            assert 2 * nv <= os.path.getsize(binf) <= 4 * nv + 4 * nl, (param, nv, os.path.getsize(binf))
            checked += 1
    assert checked >= 1


def test_damaged_models_are_rejected_with_a_message(pkg, tmp_path):
    src = parity.model_dir("rife-v4.6")
    if src is None:
        pytest.skip("no rife-v4.6 model")
    param, binf = os.path.join(src, "flownet.param"), os.path.join(src, "flownet.bin")
    text = open(param).read()
    data = open(binf, "rb").read()
    cases = {}
    cases["truncated_bin"] = (text, data[: len(data) // 2])
    cases["trailing_bytes"] = (text, data + b"\0" * 8)
    cases["bad_magic"] = (text.replace("7767517", "1234567", 1), data)
    lines = text.split("\n")
    cases["dropped_layer"] = ("\n".join(lines[:5] + lines[6:]), data)
    # a weight count of 2^31-1 in the first convolution must not turn into a multi-gigabyte allocation
    conv = next(i for i, l in enumerate(lines) if l.startswith("Convolution"))
    import re
    cases["huge_weight_count"] = ("\n".join(lines[:conv] + [re.sub(r" 6=\d+", " 6=2147483647", lines[conv])] + lines[conv + 1:]), data)
    cases["negative_weight_count"] = ("\n".join(lines[:conv] + [re.sub(r" 6=\d+", " 6=-5", lines[conv])] + lines[conv + 1:]), data)
    cases["absurd_blob_count"] = ("\n".join(lines[:conv] + [re.sub(r"^(\S+\s+\S+)\s+\d+\s+\d+", r"\1 2000000000 1", lines[conv])] + lines[conv + 1:]), data)
    cases["empty_param"] = ("", data)
    for name, (t, b) in cases.items():
        p = tmp_path / (name + ".param")
        q = tmp_path / (name + ".bin")
        p.write_text(t)
        q.write_bytes(b)
        r, _, _, _, err = _parse(pkg, str(p), str(q))
        assert r == -3 and err, (name, r, err)
    r, _, _, _, err = _parse(pkg, str(tmp_path / "missing.param"), str(tmp_path / "missing.bin"))
    assert r == -3 and "cannot open" in err
