"""The C++ host mirror builds and its CPU self-test (reference gtest known answers) passes."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "qmf_b200", "host")


def test_host_cpp_builds_and_selftest_passes():
    subprocess.check_call(["make", "-s", "-j4", "-C", HOST])
    out = subprocess.run([os.path.join(HOST, "bin", "host_selftest")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    for b in ("wals", "bpr", "gen_uniform"):
        assert os.access(os.path.join(HOST, "bin", b), os.X_OK)


def test_gen_uniform_is_seeded(tmp_path):
    a, b = str(tmp_path / "a.dat"), str(tmp_path / "b.dat")
    for f in (a, b):
        subprocess.check_call([os.path.join(HOST, "bin", "gen_uniform"), "1000", f, "42"])
    la, lb = open(a).read().split(), open(b).read().split()
    assert la == lb and len(la) == 1000 and all(abs(float(x)) <= 0.01 and len(x.split(".")[1]) == 9 for x in la)
