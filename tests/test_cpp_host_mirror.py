"""The C++ host mirror (include/zkc_b200.hpp: the reference's entry points by name, over the C ABI) compiles with g++, links the
in-tree library, refuses to start without a device, and -- on a GPU -- reproduces the oracle bit for bit from C++."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    import orc as O
    O.load()  # builds oracle/liborc.so if needed
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_test")
    lib_dir, orc_dir = os.path.join(ROOT, "era_zkevm_circuits_b200"), os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-I", orc_dir,
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", out,
                           os.path.join(lib_dir, "libzkc_b200.so"), os.path.join(orc_dir, "liborc.so"),
                           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{orc_dir}", "-ldl", "-lpthread"])
    return out


def test_compiles_links_and_refuses_without_device(binary):
    out = subprocess.run([binary, "nodevice"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "zkc_b200" in out.stdout and "25 entry points" in out.stdout
    import torch
    if not torch.cuda.is_available():
        assert "refused without a device" in out.stdout and "NO_DEVICE" in out.stdout


@pytest.mark.gpu
def test_cpp_parity_on_gpu(binary):
    out = subprocess.run([binary, "parity"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
