"""Test helper: thin torch-tensor wrappers over the single-operator entry points of the C-ABI."""
import ctypes as C

import torch

import tuber_b200  # noqa: F401  (import shim)
from tuber_b200 import _lib


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_split(x):
    rows, cols = x.shape
    out = torch.empty((rows, cols), device="cuda", dtype=torch.float32)     # same bytes as fp32
    _lib.check(_lib.load().tuber_op_to_split(P(x), P(out), rows, cols, stream()))
    return out


def from_split(s, rows, cols):
    out = torch.empty((rows, cols), device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().tuber_op_from_split(P(s), P(out), rows, cols, stream()))
    return out


def pack_weight(w):
    n, k = w.shape
    out = torch.empty((n, k), device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().tuber_op_pack_weight(P(w), P(out), n, k, stream()))
    return out


def gemm_tc(a, w, scale=None, shift=None, res=None, res_split=False, res_mod=0, relu=False, c_split=False):
    """a (M,K) fp32, w (N,K) fp32 -> (M,N) fp32 through the tcgen05 kernel (operands converted here)."""
    m, k = a.shape
    n = w.shape[0]
    a_s, w_p = to_split(a.contiguous()), pack_weight(w.contiguous())
    r = None
    if res is not None:
        r = to_split(res.contiguous()) if res_split else res.contiguous()
    c = torch.empty((m, n), device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().tuber_op_gemm_tc(P(a_s), P(w_p), P(scale), P(shift), P(r), int(res_split), res_mod, P(c),
                                            int(c_split), m, n, k, int(relu), stream()))
    return from_split(c, m, n) if c_split else c


def sgemm(a, w, bias=None, res=None, act=0):
    m, k = a.shape
    n = w.shape[0]
    c = torch.empty((m, n), device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().tuber_op_sgemm(P(a.contiguous()), P(w.contiguous()), P(bias), P(res), P(c), m, n, k, act, stream()))
    return c


def gemm_tc_fused2(a, w, scale, shift, res, w2, scale2, shift2, ab=None):
    """x' = relu(scale * ([a | ab] w^T) + shift + res) (M,N1);  t1' = relu(scale2 * (x' w2^T) + shift2) (M,N2) -- one kernel."""
    m, k = a.shape
    kb = ab.shape[1] if ab is not None else 0
    n1, n2 = w.shape[0], w2.shape[0]
    a_s, w_p, w2_p = to_split(a.contiguous()), pack_weight(w.contiguous()), pack_weight(w2.contiguous())
    ab_s = to_split(ab.contiguous()) if ab is not None else None
    r = to_split(res.contiguous()) if res is not None else None
    c = torch.empty((m, n1), device="cuda", dtype=torch.float32)
    c2 = torch.empty((m, n2), device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().tuber_op_gemm_tc_fused2(P(a_s), P(ab_s), P(w_p), P(scale), P(shift), P(r), P(c), m, k, kb, P(w2_p), P(scale2),
                                                   P(shift2), P(c2), n1, n2, stream()))
    return from_split(c, m, n1), c2
