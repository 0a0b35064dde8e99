"""Small drivers used under ncu: run one kernel family on a fixed-size problem.
   python tools/run_kernel.py atb|gemm|fc1|mixer|gru [--m ROWS]"""
import argparse
import ctypes as C
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymarl_b200 import _lib  # noqa: E402


def atb(m, c, k, reps):
    D = th.randn(m, c, device="cuda")
    A = th.randn(m, k, device="cuda")
    out = th.empty(c, k, device="cuda")
    bias = th.empty(c, device="cuda")
    need = _lib.lib().pmb_gemm_bf16_atb_workspace_bytes(m, c, k)
    scratch = th.empty(need, dtype=th.uint8, device="cuda")
    ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    for i in range(reps + 1):
        if i == 1:
            ev0.record()
        _lib.check(_lib.lib().pmb_gemm_bf16_atb(m, c, k, _lib.ptr(D), c, _lib.ptr(A), k, _lib.ptr(out), _lib.ptr(bias),
                                                _lib.ptr(scratch), need, _lib.stream_ptr()))
    ev1.record()
    th.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    gb = m * (c + k) * 4 / 1e9
    print("atb m=%d c=%d k=%d: %.3f ms  (%.1f GB/s of fp32 operands, %.1f TFLOP/s)" % (m, c, k, ms, gb / ms * 1e3, 2.0 * m * c * k / ms / 1e9))


def gemm(m, n, k, reps):
    A = th.randn(m, k, device="cuda")
    W = th.randn(n, k, device="cuda")
    out = th.empty(m, n, device="cuda")
    need = _lib.lib().pmb_gemm_bf16_workspace_bytes(n, k)
    scratch = th.empty(need, dtype=th.uint8, device="cuda")
    ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    for i in range(reps + 1):
        if i == 1:
            ev0.record()
        _lib.check(_lib.lib().pmb_gemm_bf16_tn(m, n, k, _lib.ptr(A), _lib.ptr(W), None, _lib.ptr(out), _lib.ptr(scratch),
                                               need, _lib.stream_ptr()))
    ev1.record()
    th.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print("gemm m=%d n=%d k=%d: %.3f ms  (A read %.1f GB/s, %.1f TFLOP/s)" % (m, n, k, ms, m * k * 4 / ms / 1e6, 2.0 * m * n * k / ms / 1e9))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--m", type=int, default=2_000_000)
    ap.add_argument("--c", type=int, default=192)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    if a.what == "atb":
        atb(a.m, a.c, a.k, a.reps)
    elif a.what == "gemm":
        gemm(a.m, a.n, a.k, a.reps)
