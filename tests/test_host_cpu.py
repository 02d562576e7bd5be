"""CPU tests of the C++ host classes (cwsl_digi_b200/host): built and run as a native test binary."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cwsl_digi_b200", "host")


def test_host_classes(tmp_path, cw):
    subprocess.run(["make", "-C", HOST, "all"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(HOST, "host_tests"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host tests ok" in r.stdout


def test_station_demo_fails_loudly_without_gpu(tmp_path, cw):
    if cw.device_count() > 0:
        import pytest
        pytest.skip("GPU present")
    subprocess.run(["make", "-C", HOST, "all"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(HOST, "station_demo"), os.path.join(ROOT, "tests", "data", "station_20m.ini"),
                        str(tmp_path), "1"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
