"""CPU-side checks: the C-ABI library loads and exports every symbol include/kfnet_b200.h declares,
it refuses to run without a GPU (no fallback), and the host-side records / synthetic generators behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from kfnet_b200 import _lib
    return _lib.load()


def test_header_symbols_exported_and_bound(lib):
    from kfnet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "kfnet_b200.h")).read()
    declared = set(re.findall(r"KFN_API\s+[\w\s\*]+?\b(kfn_\w+)\s*\(", hdr))
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.kfn_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    h = C.c_void_p()
    rc = lib.kfn_create(C.byref(h), 0, 0, 480, 640)
    assert rc != 0 and len(lib.kfn_last_error()) > 0
    from kfnet_b200.engine import Engine
    from kfnet_b200 import synth
    with pytest.raises(RuntimeError):
        Engine(synth.make_weights(0))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kfnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f"{f} mentions the oracle"


def test_spec_mirrors_reference_defaults():
    from kfnet_b200 import KFNetDataSpec
    s = KFNetDataSpec()
    assert (s.batch_size, s.image_size, s.scene, s.sequence_length, s.image_num) == (4, (480, 640), "stairs", 500, 2000)
    assert KFNetDataSpec(scene="fire").sequence_length == 1000 and KFNetDataSpec(scene="office").image_num == 6000


def test_synth_is_deterministic_and_shaped():
    from kfnet_b200 import synth
    w1, w2 = synth.make_weights(0), synth.make_weights(0)
    assert set(w1) == set(synth.weight_shapes())
    for k, shp in synth.weight_shapes().items():
        assert w1[k].shape == shp and w1[k].dtype == np.float32 and np.array_equal(w1[k], w2[k])
    assert w1["Temporal/conv6/kernel"].shape == (3, 3, 48, 16)        # concat(upconv0 16, conv0 32)
    assert w1["Temporal/upconv2/kernel"].shape == (3, 3, 64, 128)     # conv2d_transpose: [kh,kw,Cout,Cin]
    seq = synth.make_sequence(3, seed=1)
    assert seq.shape == (3, 480, 640, 3) and seq.dtype == np.uint8 and np.array_equal(seq, synth.make_sequence(3, seed=1))
    T = synth.make_transform(2)
    np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-5)


def test_pair_stream_and_file_io(tmp_path):
    from kfnet_b200 import eval as ev
    assert ev.full_groups_in_range(0, 4) == [[1, 0], [0, 1], [1, 2], [2, 3]]          # train.py:67-71
    p = tmp_path / "transform.txt"
    T = np.eye(4, dtype=np.float32); T[0, 3] = 2.0
    np.savetxt(p, T)
    np.testing.assert_allclose(ev.get_transform(str(p)), np.linalg.inv(T), atol=1e-6)  # train.py:49-52
    (tmp_path / "model.ckpt-5.index").write_text("")
    (tmp_path / "model.ckpt-20.index").write_text("")
    snap, step = ev.get_snapshot(str(tmp_path))
    assert step == 20 and snap.endswith("model.ckpt-20")
    lst = tmp_path / "l.txt"
    lst.write_text("a.png \nb.png\n")
    assert ev.read_lines(str(lst)) == ["a.png", "b.png"]


def test_c_consumer(tmp_path):
    """include/kfnet_b200.h is plain C: a C99 program compiles against it (-Wall -Wextra -Werror), dlopens the library,
    finds every entry point and gets loud errors (never a fallback) for bad arguments / a missing GPU."""
    import shutil
    import subprocess
    from kfnet_b200 import build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "kfn_c_smoke")
    subprocess.run([gcc, "-O1", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c_abi", "smoke.c"), "-o", exe, "-ldl"], check=True)
    r = subprocess.run([exe, build.build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c-abi smoke: ok" in r.stdout

