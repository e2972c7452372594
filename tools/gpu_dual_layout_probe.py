"""Can ONE shared-memory tile feed two UMMAs that want it in different "majors"?  A tile [128 rows r][32 floats c] with c
contiguous is MN-major (mn = c, k = r) for one product and K-major (row = r, k = c) for the other; the two canonical
128-byte-swizzle layouts only differ in the swizzle granularity (16-byte chunks ^ (r & 7) vs 32-byte chunks ^ (r & 3)).
This probe writes the tile in the MN-major SWIZZLE_128B_BASE32B image and asks the tensor core to read it as a K-major A
operand with the SAME swizzle mode (and, as a control, writes a SWIZZLE_128B image and reads it the usual way).
    python tools/gpu_dual_layout_probe.py      (needs a B200)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eeg_image_decode_b200 import _lib  # noqa: E402

L = _lib.lib()
L.eegb200_debug_umma_generic.argtypes = [ctypes.c_void_p] * 5
N = 16
rng = np.random.default_rng(0)
X = rng.integers(-8, 9, size=(128, 32)).astype(np.float32)      # tile, tf32-exact
W = rng.integers(-4, 5, size=(N, 32)).astype(np.float32)        # B operand [n][c]


def img_sw128(M, rows):                       # K-major SWIZZLE_128B: 16-byte chunk ^ (row & 7)
    out = np.zeros(8192, np.float32)
    for r in range(rows):
        for c in range(32):
            off = (r >> 3) * 1024 + (r & 7) * 128 + (((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4
            out[off // 4] = M[r, c]
    return out


def img_base32b(M, rows):                     # 32-byte chunk ^ (row & 3); one slab [128 k-rows][32 mn]
    out = np.zeros(8192, np.float32)
    for r in range(rows):
        for c in range(32):
            off = r * 128 + (((c >> 3) ^ (r & 3)) << 5) + (c & 7) * 4
            out[off // 4] = M[r, c]
    return out


def run(a_img, b_img, cfg):
    a = torch.from_numpy(a_img).cuda()
    b = torch.from_numpy(b_img).cuda()
    c = torch.tensor(cfg, dtype=torch.int32)
    out = torch.zeros(128 * cfg[10], device="cuda")
    _lib.check(L.eegb200_debug_umma_generic(_lib.ptr(a), _lib.ptr(b), c.data_ptr(), _lib.ptr(out), _lib.stream_ptr()), "probe")
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(128, cfg[10])


want = X @ W.T                                 # D[r][n] = sum_c X[r][c] W[n][c]
b_img = img_sw128(W, N)
SW128, BASE32B = 2, 1

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    # child process (a rejected descriptor kills the CUDA context): one configuration
    import json
    cfg = json.loads(sys.argv[2])
    w = eval(sys.argv[3])
    d = run(img_base32b(X, 128), b_img, cfg)
    print("OK" if np.array_equal(d, w) else "MISMATCH max %g" % np.abs(d - w).max())
    sys.exit(0)


def attempt(tag, cfg, want_):
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", json.dumps(cfg), want_], capture_output=True, text=True)
    last = (r.stdout.strip().splitlines() or ["?"])[-1]
    err = "misaligned address" if "misaligned" in r.stderr else ("illegal" if "illegal" in r.stderr else "")
    if last == "?" and not err:
        err = (r.stderr.strip().splitlines() or ["no output"])[-1][:80]
    print(f"{tag:60s}: {last if not err else 'CUDA ERROR ' + err}", flush=True)


# control: standard K-major SWIZZLE_128B A operand
d0 = run(img_sw128(X, 128), b_img, [SW128, 16, 1024, 32, 0, SW128, 16, 1024, 32, 0, N, 4])
print("control  K-major SWIZZLE_128B          :", "OK" if np.array_equal(d0, want) else f"MISMATCH max {np.abs(d0 - want).max()}")
# the BASE32B image as an MN-major B operand (the verified use): D[m][c] = sum_{r<32} V[m][r] X[r][c]
V = rng.integers(-4, 5, size=(128, 32)).astype(np.float32)
d2 = run(img_sw128(V, 128), img_base32b(X, 32), [SW128, 16, 1024, 32, 0, BASE32B, 4096, 512, 1024, 1, 32, 4])
print("control  MN-major BASE32B as B operand :", "OK" if np.array_equal(d2, V @ X[:32, :]) else "MISMATCH")
# candidate: the BASE32B image read as a K-major A operand with the BASE32B swizzle
attempt("K-major BASE32B, ONE k-step (c 0..7)", [BASE32B, 16, 1024, 32, 0, SW128, 16, 1024, 32, 0, N, 1], "X[:, :8] @ W[:, :8].T")
for lbo, sbo, kadv in ((16, 1024, 32), (16, 1024, 64), (16, 1024, 128), (1, 1024, 32), (4096, 512, 32), (16, 512, 32)):
    attempt(f"K-major BASE32B, 4 k-steps, lbo={lbo} sbo={sbo} kadv={kadv}", [BASE32B, lbo, sbo, kadv, 0, SW128, 16, 1024, 32, 0, N, 4], "want")
