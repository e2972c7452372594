"""ctypes binding of libeegdecode_b200.so (the C ABI declared in include/eegdecode_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is
raised -- the product path never routes through PyTorch eager or the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeegdecode_b200.so")
_lib = None
ABI_VERSION = 2      # include/eegdecode_b200.h EEGB200_ABI_VERSION


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
        ("round_tf32", ctypes.c_int), ("store_mode", ctypes.c_int), ("split_k", ctypes.c_int), ("tile_n", ctypes.c_int),
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
        if _lib.eegb200_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libeegdecode_b200.so ABI version mismatch (library {_lib.eegb200_abi_version()}, "
                               f"binding {ABI_VERSION}): rebuild with `python -m eeg_image_decode_b200.build`")
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
    """the CURRENT device's current stream: call inside ``on_device(...)`` so that it is the tensors' device"""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(t):
    """Context manager making the device of tensor ``t`` (or a torch.device / index) current for a library call.  The
    library launches on the current device and keeps per-device state (side streams, kernel attributes, SM count) keyed
    by it, so a model on cuda:1 works while the caller's current device is 0 (the reference's ``--gpu cuda:N`` usage)."""
    import torch
    dev = t.device if torch.is_tensor(t) else torch.device(t) if not isinstance(t, int) else torch.device("cuda", t)
    if dev.type != "cuda":
        raise RuntimeError("eegdecode_b200: expected a CUDA tensor (this path has no CPU implementation)")
    return torch.cuda.device(dev)


def launch_count() -> int:
    return int(lib().eegb200_launch_count())


def set_gemm_backend(backend: int) -> None:
    check(lib().eegb200_set_gemm_backend(int(backend)), "set_gemm_backend")


def gemm(A, B, C, M, N, K, *, lda=None, ldb=None, ldc=None, a_mn=False, b_mn=False, alpha=1.0, bias=None,
         bias_period=0, ld_bias=0, aux_out=None, ld_aux=0, act=0, drop_seed=0, drop_site=0, drop_p=0.0, drop_ld=0,
         mul_in=None, ld_mul=0, resid=None, ld_res=0, round_tf32=False, store_mode=0, split_k=1, tile_n=0):
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
    d.round_tf32, d.store_mode, d.split_k, d.tile_n = int(round_tf32), store_mode, split_k, tile_n
    with on_device(C):
        check(lib().eegb200_gemm(ctypes.byref(d), stream_ptr()), "gemm")


# ------------------------------------------------------------------------------------------------
# ATM-S / InfoNCE / retrieval / AdamW bindings
# ------------------------------------------------------------------------------------------------
P_NAMES = [  # order == enum eegb200_param; values == reference state_dict keys
    "encoder.enc_embedding.value_embedding.weight",
    "encoder.enc_embedding.value_embedding.bias",
    "encoder.enc_embedding.subject_embedding.subject_embedding.weight",
    "encoder.enc_embedding.subject_embedding.shared_embedding",
    "encoder.encoder.attn_layers.0.attention.query_projection.weight",
    "encoder.encoder.attn_layers.0.attention.query_projection.bias",
    "encoder.encoder.attn_layers.0.attention.key_projection.weight",
    "encoder.encoder.attn_layers.0.attention.key_projection.bias",
    "encoder.encoder.attn_layers.0.attention.value_projection.weight",
    "encoder.encoder.attn_layers.0.attention.value_projection.bias",
    "encoder.encoder.attn_layers.0.attention.out_projection.weight",
    "encoder.encoder.attn_layers.0.attention.out_projection.bias",
    "encoder.encoder.attn_layers.0.conv1.weight",
    "encoder.encoder.attn_layers.0.conv1.bias",
    "encoder.encoder.attn_layers.0.conv2.weight",
    "encoder.encoder.attn_layers.0.conv2.bias",
    "encoder.encoder.attn_layers.0.norm1.weight",
    "encoder.encoder.attn_layers.0.norm1.bias",
    "encoder.encoder.attn_layers.0.norm2.weight",
    "encoder.encoder.attn_layers.0.norm2.bias",
    "encoder.encoder.norm.weight",
    "encoder.encoder.norm.bias",
    "enc_eeg.0.tsconv.0.weight",
    "enc_eeg.0.tsconv.0.bias",
    "enc_eeg.0.tsconv.2.weight",
    "enc_eeg.0.tsconv.2.bias",
    "enc_eeg.0.tsconv.4.weight",
    "enc_eeg.0.tsconv.4.bias",
    "enc_eeg.0.tsconv.5.weight",
    "enc_eeg.0.tsconv.5.bias",
    "enc_eeg.0.projection.0.weight",
    "enc_eeg.0.projection.0.bias",
    "proj_eeg.0.weight",
    "proj_eeg.0.bias",
    "proj_eeg.1.fn.1.weight",
    "proj_eeg.1.fn.1.bias",
    "proj_eeg.2.weight",
    "proj_eeg.2.bias",
]
P_SUBJ_TABLE, P_SUBJ_SHARED = 2, 3
BUF_NAMES = [  # order == enum eegb200_buffer
    "encoder.enc_embedding.position_embedding.pe",
    "enc_eeg.0.tsconv.2.running_mean",
    "enc_eeg.0.tsconv.2.running_var",
    "enc_eeg.0.tsconv.5.running_mean",
    "enc_eeg.0.tsconv.5.running_var",
]
SITE_COUNT = 8
PHASE_A, PHASE_B, PHASE_C, PHASE_ALL = 1, 2, 4, 7
REF_DROPOUT_P = (0.0, 0.25, 0.25, 0.25, 0.25, 0.25, 0.5, 0.5)

PtrArrayP = ctypes.c_void_p * len(P_NAMES)
PtrArrayB = ctypes.c_void_p * len(BUF_NAMES)
FloatArrayS = ctypes.c_float * SITE_COUNT


class AtmsIO(ctypes.Structure):
    _fields_ = [
        ("params", ctypes.POINTER(ctypes.c_void_p)),
        ("buffers", ctypes.POINTER(ctypes.c_void_p)),
        ("x", ctypes.c_void_p),
        ("subject_ids", ctypes.c_void_p),
        ("B", ctypes.c_int),
        ("n_subjects", ctypes.c_int),
        ("train", ctypes.c_int),
        ("update_running_stats", ctypes.c_int),
        ("seed", ctypes.c_uint64),
        ("dropout_p", ctypes.POINTER(ctypes.c_float)),
        ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("out", ctypes.c_void_p),
        ("seed_offset_dev", ctypes.c_void_p),
        # joint-subject variant (ABI v2); joint_value_w == NULL selects the single shared value embedding
        ("joint_value_w", ctypes.POINTER(ctypes.c_void_p)),
        ("joint_value_b", ctypes.POINTER(ctypes.c_void_p)),
        ("joint_value_dw", ctypes.POINTER(ctypes.c_void_p)),
        ("joint_value_db", ctypes.POINTER(ctypes.c_void_p)),
        ("group_offsets", ctypes.POINTER(ctypes.c_int32)),
        ("group_subject", ctypes.POINTER(ctypes.c_int32)),
        ("n_groups", ctypes.c_int),
    ]


class InfoNceIO(ctypes.Structure):
    _fields_ = [
        ("eeg", ctypes.c_void_p), ("tgt_img", ctypes.c_void_p), ("tgt_txt", ctypes.c_void_p),
        ("B", ctypes.c_int), ("N", ctypes.c_int), ("D", ctypes.c_int), ("row_offset", ctypes.c_int),
        ("logit_scale", ctypes.c_void_p),
        ("w_img", ctypes.c_float), ("w_txt", ctypes.c_float), ("grad_out", ctypes.c_float),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("col_stats", ctypes.c_void_p), ("col_parts", ctypes.c_void_p), ("n_parts", ctypes.c_int),
        ("loss", ctypes.c_void_p), ("d_eeg", ctypes.c_void_p), ("d_logit_scale", ctypes.c_void_p),
    ]


def _sig():
    L = lib()
    if getattr(L, "_eeg_sig_done", False):
        return L
    L.eegb200_atms_workspace_bytes.restype = ctypes.c_size_t
    L.eegb200_atms_workspace_bytes.argtypes = [ctypes.c_int]
    L.eegb200_infonce_workspace_bytes.restype = ctypes.c_size_t
    L.eegb200_infonce_workspace_bytes.argtypes = [ctypes.c_int] * 4
    L.eegb200_atms_forward.argtypes = [ctypes.POINTER(AtmsIO), ctypes.c_int, ctypes.c_void_p]
    L.eegb200_atms_backward.argtypes = [ctypes.POINTER(AtmsIO), ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                         ctypes.c_int, ctypes.c_void_p]
    L.eegb200_atms_ws_tensor.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int)]
    L.eegb200_infonce.argtypes = [ctypes.POINTER(InfoNceIO), ctypes.c_int, ctypes.c_void_p]
    L.eegb200_retrieval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]
    L.eegb200_adamw_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_longlong, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    L.eegb200_adamw_step_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_longlong, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    L.eegb200_mse.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_float,
                              ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.eegb200_dropout_mask.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L._eeg_sig_done = True
    return L


def atms_workspace_bytes(B: int) -> int:
    return int(_sig().eegb200_atms_workspace_bytes(int(B)))


def infonce_workspace_bytes(B: int, N: int, D: int, nt: int) -> int:
    return int(_sig().eegb200_infonce_workspace_bytes(int(B), int(N), int(D), int(nt)))


def infonce_target_offset(B: int, N: int, D: int, nt: int) -> int:
    L = _sig()
    L.eegb200_infonce_target_offset.restype = ctypes.c_size_t
    L.eegb200_infonce_target_offset.argtypes = [ctypes.c_int] * 4
    return int(L.eegb200_infonce_target_offset(int(B), int(N), int(D), int(nt)))


def tf32_round(src, dst) -> None:
    """dst = src rounded to TF32 (round to nearest; plain copy when EEGB200_TF32_ROUND=0), [rows, D] contiguous fp32"""
    with on_device(src):
        check(_sig().eegb200_tf32_round(ptr(src), ptr(dst), int(src.shape[0]), int(src.shape[1]), stream_ptr()), "tf32_round")


def peer_sum_buffer_bytes() -> int:
    L = lib()
    L.eegb200_peer_sum_buffer_bytes.restype = ctypes.c_size_t
    return int(L.eegb200_peer_sum_buffer_bytes())


def peer_sum_f64(data, peer_ptr_array, rank: int, world: int, seq_dev, err_dev) -> None:
    """in-place one-shot all-reduce of a small float64 tensor over peer memory (csrc/peer_sum.cu)"""
    with on_device(data):
        check(lib().eegb200_peer_sum_f64(ptr(data), int(data.numel()), peer_ptr_array, int(rank), int(world), ptr(seq_dev),
                                         ptr(err_dev), stream_ptr()), "peer_sum_f64")


def atms_forward(io: AtmsIO, phases: int, device) -> None:
    with on_device(device):
        check(_sig().eegb200_atms_forward(ctypes.byref(io), phases, stream_ptr()), "atms_forward")


def atms_backward(io: AtmsIO, d_out, grads_array, phases: int, device) -> None:
    with on_device(device):
        check(_sig().eegb200_atms_backward(ctypes.byref(io), ptr(d_out), grads_array, phases, stream_ptr()), "atms_backward")


def ws_tensor(workspace, B: int, name: str):
    """view of a named intermediate inside an ATM-S workspace (float32, or float64 for the *_sums entries)"""
    import torch
    p, r, c, ld = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(_sig().eegb200_atms_ws_tensor(ptr(workspace), B, name.encode(), ctypes.byref(p), ctypes.byref(r),
                                        ctypes.byref(c), ctypes.byref(ld)), "ws_tensor")
    off = p.value - workspace.data_ptr()
    if name.endswith("_sums"):
        return workspace[off:off + 80 * 8].view(torch.float64)
    n = r.value * ld.value
    return workspace[off:off + n * 4].view(torch.float32).view(r.value, ld.value)[:, :c.value]


def infonce(io: InfoNceIO, phases: int, device) -> None:
    with on_device(device):
        check(_sig().eegb200_infonce(ctypes.byref(io), phases, stream_ptr()), "infonce")


def adamw_step(p, g, m, v, n, lr, b1, b2, eps, wd, step) -> None:
    with on_device(p):
        check(_sig().eegb200_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), int(n), lr, b1, b2, eps, wd, int(step),
                                        stream_ptr()), "adamw_step")


def adamw_step_dev(p, g, m, v, n, lr, b1, b2, eps, wd, step_dev) -> None:
    with on_device(p):
        check(_sig().eegb200_adamw_step_dev(ptr(p), ptr(g), ptr(m), ptr(v), int(n), lr, b1, b2, eps, wd, ptr(step_dev),
                                            stream_ptr()), "adamw_step_dev")


def mse(eeg, tgt, n_total_rows: int, weight: float, grad_out: float, loss=None, loss_term=None, d_eeg=None) -> None:
    """weight * MSE(eeg, tgt) share of these rows (mean over n_total_rows*D elements): loss / loss_term (1-element device
    tensors) += ; d_eeg += weight*grad_out * dMSE/deeg"""
    B, D = eeg.shape
    with on_device(eeg):
        check(_sig().eegb200_mse(ptr(eeg), ptr(tgt), B, D, int(n_total_rows), float(weight), float(grad_out), ptr(loss),
                                 ptr(loss_term), ptr(d_eeg), stream_ptr()), "mse")


def l2norm_forward(x, y, norms) -> None:
    with on_device(x):
        check(lib().eegb200_l2norm_forward(ptr(x), ptr(y), ptr(norms), int(x.shape[0]), int(x.shape[1]), stream_ptr()),
              "l2norm_forward")


def l2norm_backward(y, norms, dy, dx) -> None:
    with on_device(y):
        check(lib().eegb200_l2norm_backward(ptr(y), ptr(norms), ptr(dy), ptr(dx), int(y.shape[0]), int(y.shape[1]),
                                            stream_ptr()), "l2norm_backward")


def dropout_mask(seed: int, site: int, p: float, rows: int, cols: int, ld: int, device="cuda"):
    import torch
    out = torch.empty(rows, cols, device=device, dtype=torch.float32)
    with on_device(out):
        check(_sig().eegb200_dropout_mask(seed, site, p, rows, cols, ld, ptr(out), stream_ptr()), "dropout_mask")
    return out


def retrieval(eeg, gallery, logit_scale, sel=None, labels=None, want_top5=True):
    """scores = logit_scale * eeg @ gallery.T on device; returns dict(top1, top5, correct, logits)."""
    import torch
    Q, D = eeg.shape
    G = gallery.shape[0]
    ld = (G + 3) // 4 * 4
    dev = eeg.device
    logits = torch.empty(Q, ld, device=dev, dtype=torch.float32)
    round_ws = torch.empty(3 * (Q + G) * D, device=dev, dtype=torch.float32)
    top1 = torch.empty(Q, device=dev, dtype=torch.int64)
    top5 = torch.empty(Q, 5, device=dev, dtype=torch.int32) if want_top5 else None
    correct = torch.zeros(1, device=dev, dtype=torch.int32)
    k = 0
    sel_ws = None
    if sel is not None:
        sel = sel.to(device=dev, dtype=torch.int32).contiguous()
        k = sel.shape[1]
        sel_ws = torch.empty(Q, k, device=dev, dtype=torch.float32)
    if labels is not None:
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    with on_device(eeg):
        check(_sig().eegb200_retrieval(ptr(eeg.contiguous()), ptr(gallery.contiguous()), Q, G, D, ptr(logit_scale),
                                       ptr(logits), ld, ptr(round_ws), ptr(sel), k, ptr(sel_ws), ptr(labels), ptr(correct),
                                       ptr(top1), ptr(top5), stream_ptr()), "retrieval")
    return {"top1": top1, "top5": top5, "correct": correct, "logits": logits[:, :G]}


def prof_enable(on: bool) -> None:
    lib().eegb200_prof_enable(int(on))


def prof_report() -> dict:
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().eegb200_prof_report(buf, ctypes.c_size_t(len(buf))), "prof_report")
    return json.loads(buf.value.decode())
