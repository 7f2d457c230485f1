"""The CTA-pair GEMM kernel (cta_group::2, csrc/gemm.cu conv_gemm_kernel<true>) must give BIT-IDENTICAL results to the
one-CTA kernel: same K order, same fp32 accumulation in TMEM, same epilogue -- only who fetches which half of B
changes.  The variant is an environment switch read once per process, so each one runs in its own process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    out = subprocess.run([sys.executable, os.path.join(HERE, "pair_digest.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return dict(line.split() for line in out.stdout.splitlines() if line.strip())


def test_pair_kernel_is_bit_identical_to_single_cta_kernel(cuda):
    single = _run({"SONIC_GEMM_PAIR": "0"})
    pair = _run({"SONIC_GEMM_PAIR": "1", "SONIC_GEMM_PAIR_MIN": "0"})
    assert single.keys() == pair.keys() and len(single) >= 14
    diff = [name for name in single if single[name] != pair[name]]
    assert not diff, diff
