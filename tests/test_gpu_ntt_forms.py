"""The NTT butterflies have two multipliers (csrc/ntt.cu): fixed-operand (Shoup) twiddle pairs — the default — and
Montgomery-form twiddles with the word-serial product (EON_NTT_SHOUP=0).  The choice is read once per process, so
the DFT parity suite is run a second time in a child process with the other form: both must reproduce the
reference's canonical limbs (dft/src/traits.rs:61-249 semantics, tests/test_gpu_dft.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dft_parity_with_montgomery_twiddles():
    from plonky3_eon_b200 import lib
    assert int(lib.load().eon_ntt_twiddle_form()) == int(os.environ.get("EON_NTT_SHOUP", "1") != "0")
    env = dict(os.environ, EON_NTT_SHOUP="0")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_dft.py"), "-q", "-x",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["1", "2"])
def test_dft_parity_with_double_buffered_passes(mode):
    """EON_NTT_DB selects the persistent double-buffered pass kernel (k_ntt_pass_db, off by default: measured
    slower): same butterflies, so the same canonical limbs."""
    env = dict(os.environ, EON_NTT_DB=mode)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_dft.py"), "-q", "-x",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
