"""GPU tier: the C++ host-side mirror (include/qrl_b200_gr.hpp: make_gr_demod_* / make_gr_mod_* / gr_bit_sink)
built with g++ against libqrl_b200.so and run end to end (TX -> RX loop-back through the sinks)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_adapter_test(tmpdir):
    exe = os.path.join(str(tmpdir), "test_adapter")
    libdir = os.path.join(ROOT, "qradiolink_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp"), "-o", exe,
                           "-L" + libdir, "-lqrl_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_adapter_compiles(tmp_path):
    """CPU tier part: the header-only adapter and its test program compile and link against the C ABI."""
    assert os.path.exists(build_adapter_test(tmp_path))


@pytest.mark.gpu
def test_cpp_adapter_loopback(tmp_path):
    exe = build_adapter_test(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
