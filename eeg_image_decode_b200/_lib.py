"""ctypes binding of libeegdecode_b200.so (the C ABI declared in include/eegdecode_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is
raised -- the product path never routes through PyTorch eager or the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeegdecode_b200.so")
_lib = None


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int), ("a_mn_major", ctypes.c_int),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int), ("b_mn_major", ctypes.c_int),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_int),
        ("alpha", ctypes.c_float),
        ("bias", ctypes.c_void_p), ("bias_period", ctypes.c_int), ("ld_bias", ctypes.c_int),
        ("aux_out", ctypes.c_void_p), ("ld_aux", ctypes.c_int),
        ("act", ctypes.c_int),
        ("drop_seed", ctypes.c_uint64), ("drop_site", ctypes.c_uint32), ("drop_p", ctypes.c_float),
        ("drop_ld", ctypes.c_int),
        ("mul_in", ctypes.c_void_p), ("ld_mul", ctypes.c_int),
        ("resid", ctypes.c_void_p), ("ld_res", ctypes.c_int),
        ("round_tf32", ctypes.c_int), ("store_mode", ctypes.c_int), ("split_k", ctypes.c_int),
    ]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m eeg_image_decode_b200.build` "
                "(there is no CPU/eager fallback for this path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.eegb200_last_error.restype = ctypes.c_char_p
        _lib.eegb200_launch_count.restype = ctypes.c_longlong
        if _lib.eegb200_abi_version() != 1:
            raise RuntimeError("libeegdecode_b200.so ABI version mismatch")
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().eegb200_last_error().decode(errors="replace")
        raise RuntimeError(f"eegdecode_b200 {what} failed (rc={rc}): {msg}")


def ptr(t) -> ctypes.c_void_p:
    """device pointer of a CUDA tensor (None -> NULL)"""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("eegdecode_b200: expected a CUDA tensor (this path has no CPU implementation)")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().eegb200_launch_count())


def set_gemm_backend(backend: int) -> None:
    check(lib().eegb200_set_gemm_backend(int(backend)), "set_gemm_backend")


def gemm(A, B, C, M, N, K, *, lda=None, ldb=None, ldc=None, a_mn=False, b_mn=False, alpha=1.0, bias=None,
         bias_period=0, ld_bias=0, aux_out=None, ld_aux=0, act=0, drop_seed=0, drop_site=0, drop_p=0.0, drop_ld=0,
         mul_in=None, ld_mul=0, resid=None, ld_res=0, round_tf32=False, store_mode=0, split_k=1):
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = ptr(A), int(lda if lda is not None else A.stride(0)), int(a_mn)
    d.B, d.ldb, d.b_mn_major = ptr(B), int(ldb if ldb is not None else B.stride(0)), int(b_mn)
    d.C, d.ldc = ptr(C), int(ldc if ldc is not None else C.stride(0))
    d.alpha = alpha
    d.bias, d.bias_period, d.ld_bias = ptr(bias), bias_period, ld_bias
    d.aux_out, d.ld_aux = ptr(aux_out), ld_aux
    d.act = act
    d.drop_seed, d.drop_site, d.drop_p, d.drop_ld = drop_seed, drop_site, drop_p, drop_ld
    d.mul_in, d.ld_mul = ptr(mul_in), ld_mul
    d.resid, d.ld_res = ptr(resid), ld_res
    d.round_tf32, d.store_mode, d.split_k = int(round_tf32), store_mode, split_k
    check(lib().eegb200_gemm(ctypes.byref(d), stream_ptr()), "gemm")
