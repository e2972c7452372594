"""Bring-up probe for the tcgen05 GEMM: each configuration runs in its own subprocess (a trapped
kernel poisons the CUDA context) with a timeout; prints error statistics against torch fp64."""
import json
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # M, N, K, a_mn, b_mn, split_k
    (128, 64, 32, 0, 0, 1),
    (128, 256, 256, 0, 0, 1),
    (256, 128, 64, 0, 0, 1),
    (300, 250, 250, 0, 0, 1),
    (1024, 768, 256, 0, 0, 1),
    (128, 64, 32, 0, 1, 1),
    (256, 256, 256, 0, 1, 1),
    (128, 64, 32, 1, 0, 1),
    (256, 256, 256, 1, 0, 1),
    (128, 64, 32, 1, 1, 1),
    (256, 256, 2048, 1, 1, 4),
    (40, 2520, 1152, 1, 1, 3),
    (1152, 40, 2520, 0, 0, 1),
    (4096, 1024, 1440, 0, 0, 1),
]


def child(case):
    import torch
    sys.path.insert(0, ROOT)
    from eeg_image_decode_b200 import _lib
    M, N, K, a_mn, b_mn, split = case
    g = torch.Generator(device="cuda").manual_seed(1)
    pad = lambda x: (x + 3) // 4 * 4
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    ref = (A.double() @ B.double().T)
    As = torch.zeros(K, pad(M), device="cuda") if a_mn else torch.zeros(M, pad(K), device="cuda")
    if a_mn: As[:, :M] = A.T
    else: As[:, :K] = A
    Bs = torch.zeros(K, pad(N), device="cuda") if b_mn else torch.zeros(N, pad(K), device="cuda")
    if b_mn: Bs[:, :N] = B.T
    else: Bs[:, :K] = B
    out = {}
    for backend in (1, 0):
        _lib.set_gemm_backend(backend)
        C = torch.zeros(M, pad(N), device="cuda")
        _lib.gemm(As, Bs, C, M, N, K, a_mn=a_mn, b_mn=b_mn, store_mode=2 if split > 1 else 0, split_k=split)
        torch.cuda.synchronize()
        err = (C[:, :N].double() - ref).abs().max().item()
        out["simt" if backend else "tc"] = err / ref.abs().max().item()
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(json.loads(sys.argv[1]))
        sys.exit(0)
    for case in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, json.dumps(case)], capture_output=True, text=True, timeout=120)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            tail = (r.stderr.strip().splitlines() or [""])[-1][:200]
            print(case, lines[-1] if lines else f"NO RESULT rc={r.returncode} {tail}", flush=True)
            if not lines:
                print("   stdout:", r.stdout[-300:].replace("\n", " | "))
        except subprocess.TimeoutExpired:
            print(case, "TIMEOUT", flush=True)
