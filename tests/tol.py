"""Tolerance selection for the GPU parity tests.

The default build stores tensor-core operands as bf16 (the fast path the bench measures); its asserted bars are the measured
bf16 operand-rounding error with margin.  The IEEE-half operand build (`DCPT_OPERAND=fp16`, libdcpt_sm100_fp16.so: 8x smaller
operand rounding) is the parity mode: there every network-level FORWARD check asserts the north-star's 1e-3 relative bar, and
the backward bars are the measured fp16 figures with margin.  tests/test_gpu_parity_fp16.py re-runs the GPU suites in that mode.
"""
import os

FP16 = os.getenv("DCPT_OPERAND", "bf16").lower() == "fp16"
NORTH_STAR = 1e-3


def tol(bf16_tol, fp16_tol=None):
    """bf16_tol for the default build; fp16_tol (default: the north-star 1e-3) for the parity build."""
    if FP16:
        return NORTH_STAR if fp16_tol is None else fp16_tol
    return bf16_tol


def report(what, **vals):
    print(f"[parity:{'fp16' if FP16 else 'bf16'}] {what}: " + ", ".join(f"{k}={v:.3e}" for k, v in vals.items()), flush=True)
