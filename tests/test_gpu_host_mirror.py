"""The C++ mirror of the reference eval API (csrc/host/nnue_state.h) driven like the engine:
datagen form, search form with push/pop and lazy evaluation, and the batched form."""
import os
import subprocess
import sys

import pytest

from stormphrax_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_nnue_state.cpp")


def _compile(tmp_path) -> str:
    lib = build.build()
    exe = str(tmp_path / "test_nnue_state")
    subprocess.run(
        ["g++", "-std=c++17", "-O2", "-pthread", "-o", exe, SRC, f"-L{os.path.dirname(lib)}", "-lsp_nnue", f"-Wl,-rpath,{os.path.dirname(lib)}"],
        check=True,
    )
    return exe


def test_host_mirror_compiles_and_links(tmp_path):
    """CPU box: the mirror's symbols are in the library and the engine-style driver links."""
    _compile(tmp_path)


@pytest.mark.gpu
def test_host_mirror_engine_protocol(tmp_path, net):
    exe = _compile(tmp_path)
    path = tmp_path / "synthetic.nnue"
    net.image.tofile(path)
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=600)
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert "0 failures" in r.stdout
