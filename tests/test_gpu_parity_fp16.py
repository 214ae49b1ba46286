"""The GPU parity suites once more, against the IEEE-half operand build (DCPT_OPERAND=fp16 -> libdcpt_sm100_fp16.so).

VERDICT r1 "What's weak" #1 / "Next" #3: the north-star bar is 1e-3 relative to the fp32 reference; the bf16-operand fast path
sits at 1-2e-3 on 36-block stacks (the reference itself under bf16 autocast: 2.05e-3).  The parity build keeps
every kernel and layout and only changes the 16-bit operand format (8 -> 11 significand bits); in that mode tests/tol.py makes
every network-level forward check assert 1e-3 and the backward checks assert the measured fp16 figures.  The library is
process-wide, so the suites run in a child interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = ["tests/test_gpu_kernels.py", "tests/test_gpu_nafnet.py", "tests/test_gpu_restormer.py", "tests/test_gpu_dchead.py",
          "tests/test_gpu_promptir.py", "tests/test_gpu_models.py"]


@pytest.mark.skipif(os.getenv("DCPT_OPERAND", "bf16").lower() == "fp16", reason="already inside the fp16 child run")
def test_gpu_suites_with_fp16_operands():
    env = dict(os.environ, DCPT_OPERAND="fp16")
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-s", "-p", "no:cacheprovider"] + SUITES,
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    lines = r.stdout.splitlines()
    keep = ("[parity:fp16]", "TransformerBlock", "Restormer", "white-noise", "DCPT step", "grad err", "nafnet grads", "w64 ")
    print("\n".join(l for l in lines if l.startswith(keep)))
    print("\n".join(lines[-15:]))
    assert r.returncode == 0, "fp16-operand run failed:\n" + "\n".join(l for l in lines if l.startswith(keep + ("E  ", "FAILED", "tests/")))[-6000:]
