"""The device assertions that have not run on a B200 yet (tests/test_zz_gpu_late.py), exercised end to end -- host glue
and kernels -- against the whole product library compiled for the host (tests/host_emul/fullhost: stand-in CUDA
runtime, one fiber per CUDA thread).  A subprocess, because the redirection of the ctypes binding
(PHB200_TEST_HOST_EMUL=1, tests/conftest.py) is per test process; the build takes about a minute the first time."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_late_device_assertions_pass_under_host_emulation():
    env = dict(os.environ, PHB200_TEST_HOST_EMUL="1")
    with tempfile.TemporaryFile("w+") as log:
        p = subprocess.Popen([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_zz_gpu_late.py"), "-q",
                              "-m", "gpu", "-p", "no:cacheprovider"], cwd=ROOT, env=env, stdout=log,
                             stderr=subprocess.STDOUT, stdin=subprocess.DEVNULL)
        try:
            rc = p.wait(timeout=900)
        except subprocess.TimeoutExpired:
            p.kill()
            rc = -9
        log.seek(0)
        tail = log.read()[-3000:]
    assert rc == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
