"""The C++ header shim (cvgpuspeedup_b200/include/cvGPUSpeedup.cuh): compiles against the OpenCV stand-in with
plain g++ (CPU check) and passes the reference's restated test programs on a GPU (tests/cpp/test_shim.cpp)."""
import os
import subprocess
import textwrap

import pytest

from tests import util

CPP = os.path.join(util.ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "_build", "test_shim")


def _build():
    util.oracle_lib()  # makes sure oracle/liboracle.so exists
    res = subprocess.run(["make", "-C", CPP], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]


def test_shim_compiles_and_links():
    _build()
    assert os.path.exists(BIN)


def test_shim_rejects_chains_outside_the_hot_path(tmp_path):
    """Unsupported instantiations fail at compile time with a message naming the restriction."""
    src = tmp_path / "bad.cpp"
    src.write_text(textwrap.dedent(f"""
        #include "{util.ROOT}/cvgpuspeedup_b200/include/cvGPUSpeedup.cuh"
        int main() {{
            std::array<cv::cuda::GpuMat, 2> crops;
            auto r = cvGS::resize<CV_32FC3, cv::INTER_LINEAR, 2>(crops, cv::Size(8, 8), 2);
            return 0;
        }}"""))
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-DCVGS_FORCE_OPENCV_DOUBLE", "-I/usr/local/cuda/include",
                          "-x", "c++", str(src)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode != 0 and "sources with 3 or 4 channels" in res.stdout


@pytest.mark.gpu
def test_shim_runs_reference_test_programs():
    if not os.path.exists(BIN):
        _build()
    res = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "all passed" in res.stdout, res.stdout[-3000:]
