"""C++-level check of the host RNG building blocks (fetching_b200/csrc/host/rng.hpp): shared replay stream and
cached Box-Muller products against std::mt19937 and the scalar normal_distribution restatement.  CPU only."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_rng_building_blocks(tmp_path):
    exe = str(tmp_path / "host_rng_check")
    src = os.path.join(ROOT, "tests", "cpp", "host_rng_check.cpp")
    inc = os.path.join(ROOT, "fetching_b200", "csrc", "host")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-I" + inc, src, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host rng check: ok" in out.stdout
