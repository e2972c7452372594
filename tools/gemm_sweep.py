"""Micro-benchmark of the step's GEMM shapes over (tile_n, split_k): CUDA-event time per launch, L2 flushed between
launches by rotating over operand copies larger than L2 where the operands are big."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from eeg_image_decode_b200 import _lib

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def bench(M, N, K, a_mn, b_mn, split, tile_n, reps=6):
    lda = ((M if a_mn else K) + 3) // 4 * 4
    ldb = ((N if b_mn else K) + 3) // 4 * 4
    A = torch.randn((K if a_mn else M), lda, device=dev)
    Bm = torch.randn((K if b_mn else N), ldb, device=dev)
    ldc = (N + 3) // 4 * 4
    C = torch.zeros(M, ldc, device=dev)
    ts = []
    for r in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.gemm(A, Bm, C, M, N, K, lda=lda, ldb=ldb, ldc=ldc, a_mn=a_mn, b_mn=b_mn, store_mode=2 if split > 1 else 0,
                  split_k=split, tile_n=tile_n)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[1]


CASES = [
    ("retrieval 3xTF32", 1024, 1654, 3072, 0, 0, [(1, 0), (1, 64), (1, 128), (2, 128), (1, 256), (2, 256), (3, 256), (4, 256)]),
    ("logits", 1024, 2048, 1024, 0, 0, [(1, 0), (1, 128), (1, 256), (2, 256), (2, 128)]),
    ("dE", 1024, 1024, 2048, 0, 1, [(1, 0), (1, 128), (2, 128), (1, 256), (2, 256), (4, 256)]),
    ("proj1 fwd", 1024, 1024, 1440, 0, 0, [(1, 0), (1, 128), (2, 128), (2, 256), (4, 256)]),
    ("proj2 fwd", 1024, 1024, 1024, 0, 0, [(1, 0), (1, 128), (2, 128), (2, 256), (4, 256)]),
    ("dfeat", 1024, 1440, 1024, 0, 1, [(1, 0), (1, 128), (2, 128), (2, 256)]),
    ("dWp1", 1024, 1440, 1024, 1, 1, [(3, 0), (2, 0), (4, 0), (2, 256), (3, 256), (2, 128), (1, 128)]),
    ("dWp2", 1024, 1024, 1024, 1, 1, [(5, 0), (4, 0), (2, 0), (4, 256), (2, 128), (1, 128)]),
    ("dWs", 40, 2520, 36864, 1, 1, [(15, 0), (14, 0), (7, 0), (28, 0), (14, 128), (7, 128)]),
    ("dWqkv", 768, 256, 65536, 1, 1, [(25, 0), (24, 0), (12, 0), (48, 0), (12, 128), (24, 128)]),
    ("dW 256x256", 256, 256, 65536, 1, 1, [(74, 0), (37, 0), (148, 0), (37, 128), (74, 128)]),
    ("spatial fwd", 36864, 40, 2520, 0, 0, [(1, 0)]),
    ("QKV fwd", 65536, 768, 256, 0, 0, [(1, 0), (1, 128)]),
    ("token 256", 65536, 256, 256, 0, 0, [(1, 0), (1, 128)]),
    ("dH0", 65536, 256, 768, 0, 1, [(1, 0), (1, 128)]),
]
only = sys.argv[1:] 
for name, M, N, K, amn, bmn, cfgs in CASES:
    if only and not any(o in name for o in only):
        continue
    out = []
    for split, tn in cfgs:
        try:
            out.append(f"s{split}/t{tn}: {bench(M, N, K, amn, bmn, split, tn):6.1f}")
        except Exception as ex:
            out.append(f"s{split}/t{tn}: ERR {str(ex)[:40]}")
    print(f"{name:18s} M={M} N={N} K={K} {'mn' if amn else 'k'}{'n' if bmn else 'k'} | " + " | ".join(out), flush=True)
