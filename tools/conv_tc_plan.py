"""Executable restatement (CPU, torch) of the round-2 plan for the PatchEmbedding conv stack on tcgen05
(docs/ROUND2_CONV_TCGEN05.md): same blocking as the planned kernels -- 3-sample / 128-row tiles, per-channel streaming
of the spatial contraction, K padded to 32 / 64, hi/lo TF32 split of the conv operands, prefix-scan pooling, folded
BatchNorm-backward constants, channel-major weight-gradient accumulation -- checked against autograd through the plain
torch ops of the reference (ATMS_retrieval.py:101-106).  Not product code: it exists so that every tile of the CUDA
kernels can be compared against a known-good host value while they are being brought up.

    python tools/conv_tc_plan.py            # prints the max errors of every stage, exits non-zero on a mismatch
"""
import sys

import torch
import torch.nn.functional as F

N_CH, N_T, N_FILT, K_TEMP, K_POOL, S_POOL, N_POOL, N_PSUM = 63, 250, 40, 25, 51, 5, 36, 200
TILE_S = 3                      # samples per 128-row UMMA tile (3 x 36 = 108 valid rows)
EPS = 1e-5


def tf32_rn(x):
    """cvt.rna.tf32.f32: round to nearest, ties away, 10-bit mantissa"""
    u = x.contiguous().view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF
    return u.view(torch.float32)


def split_hi_lo(x):
    hi = tf32_rn(x)
    return hi, tf32_rn(x - hi)


def mma(a, b):
    """fp32-accumulated product of TF32 operands: a [M,K] . b[N,K]^T"""
    return a.double().matmul(b.double().T).float()


# ------------------------------------------------------------------------------------------------ shared pieces
def pooled_sums(x_row):
    """ps[u] = sum_{v<51} x[u+v] for u < 200 from an exclusive prefix sum (what a warp scan produces)"""
    c = torch.zeros(257)
    c[1:251] = torch.cumsum(x_row[:250], 0)
    c[251:] = c[250]
    u = torch.arange(N_PSUM)
    return c[u + K_POOL] - c[u]


def im2col_tile(x3, b0, c):
    """A operand of the conv MMA for samples b0..b0+2, channel c: [128, 32], row s*36+p = ps[5p .. 5p+24], zero padded"""
    A = torch.zeros(128, 32)
    for s in range(TILE_S):
        if b0 + s >= x3.shape[0]:
            break
        ps = pooled_sums(x3[b0 + s, c])
        for p in range(N_POOL):
            A[s * N_POOL + p, :K_TEMP] = ps[5 * p:5 * p + K_TEMP]
    return A


def conv_weights(wt):
    """B operand of the conv MMA: [48, 32] = wt[k, t] / 51, zero padded"""
    Bm = torch.zeros(48, 32)
    Bm[:N_FILT, :K_TEMP] = wt.reshape(N_FILT, K_TEMP) / K_POOL
    return Bm


def conv_tile(A, Bm_hi, Bm_lo, bt):
    """3xTF32 conv MMA of one tile: [128, 40] = A . B^T + bias"""
    a_hi, a_lo = split_hi_lo(A)
    acc = mma(a_lo, Bm_hi) + mma(a_hi, Bm_lo) + mma(a_hi, Bm_hi)
    return acc[:, :N_FILT] + bt[None, :]


def elu(z):
    return torch.where(z > 0, z, torch.expm1(z))


def elu_grad(z):
    return torch.where(z > 0, torch.ones_like(z), torch.exp(z))


# ------------------------------------------------------------------------------------------------ forward
def forward_plan(x3, P):
    """F1 + F2.  x3 [B, 63, 250].  Returns y2 [B, 36, 40] (rows (b,p)), BN1 batch statistics and per-row valid masks."""
    B = x3.shape[0]
    Bc = conv_weights(P["wt"])
    Bc_hi, Bc_lo = split_hi_lo(Bc)
    n = B * N_CH * N_POOL
    # ---- F1: statistics only ----
    s1 = torch.zeros(N_FILT, dtype=torch.float64)
    s2 = torch.zeros(N_FILT, dtype=torch.float64)
    for b0 in range(0, B, TILE_S):
        rows = min(TILE_S, B - b0) * N_POOL
        for c in range(N_CH):
            y = conv_tile(im2col_tile(x3, b0, c), Bc_hi, Bc_lo, P["bt"])[:rows]
            s1 += y.double().sum(0)
            s2 += (y.double() ** 2).sum(0)
    mean = (s1 / n).float()
    var = (s2 / n - (s1 / n) ** 2).float()
    rstd = torch.rsqrt(var + EPS)
    # ---- F2: apply, spatial contraction accumulated over the channels ----
    sc = P["g1"] * rstd
    sh = P["b1"] - mean * sc
    # spatial weights of channel c as B operand [48, 64]: Ws_c[j, k] = Ws[j, k, c], K padded 40 -> 64
    Ws = P["ws"].reshape(N_FILT, N_FILT, N_CH)
    y2 = torch.zeros(B, N_POOL, N_FILT)
    for b0 in range(0, B, TILE_S):
        ns = min(TILE_S, B - b0)
        acc = torch.zeros(128, 48)
        for c in range(N_CH):
            y = conv_tile(im2col_tile(x3, b0, c), Bc_hi, Bc_lo, P["bt"])
            a1 = torch.zeros(128, 64)
            a1[:, :N_FILT] = tf32_rn(elu(y * sc[None, :] + sh[None, :]))
            a1[ns * N_POOL:] = 0.0                                   # pad rows
            Wc = torch.zeros(48, 64)
            Wc[:N_FILT, :N_FILT] = tf32_rn(Ws[:, :, c])
            acc += mma(a1, Wc)
        y2[b0:b0 + ns] = (acc[:ns * N_POOL, :N_FILT] + P["bs"][None, :]).reshape(ns, N_POOL, N_FILT)
    return y2, mean, var, rstd


# ------------------------------------------------------------------------------------------------ backward
def backward_plan(x3, P, dy2, mean, rstd):
    """B1 + B2.  dy2 [B, 36, 40] = d loss / d y2 (rows (b,p)).  Returns dict of gradients."""
    B = x3.shape[0]
    n = B * N_CH * N_POOL
    Bc = conv_weights(P["wt"])
    Bc_hi, Bc_lo = split_hi_lo(Bc)
    Ws = P["ws"].reshape(N_FILT, N_FILT, N_CH)
    dy2_t = tf32_rn(dy2)
    sc = P["g1"] * rstd
    sh = P["b1"] - mean * sc

    def tile_dy2(b0):
        t = torch.zeros(128, 48)
        ns = min(TILE_S, B - b0)
        t[:ns * N_POOL, :N_FILT] = dy2_t[b0:b0 + ns].reshape(ns * N_POOL, N_FILT)
        return t, ns

    def tile_quantities(b0, c):
        dyt, ns = tile_dy2(b0)
        # dA1[(b,p), k] = sum_j dY2[(b,p), j] Ws[j, k, c]: A = dY2 tile [128, 48(j)], B = Ws_c^T [48(k), 48(j)]
        WcT = torch.zeros(48, 48)
        WcT[:N_FILT, :N_FILT] = tf32_rn(Ws[:, :, c]).T
        da1 = mma(dyt, WcT)[:, :N_FILT]
        A = im2col_tile(x3, b0, c)
        y = conv_tile(A, Bc_hi, Bc_lo, P["bt"])
        yhat = (y - mean[None, :]) * rstd[None, :]
        z = y * sc[None, :] + sh[None, :]
        dz = da1 * elu_grad(z)
        valid = torch.zeros(128, 1)
        valid[:ns * N_POOL] = 1.0
        return dyt, A, y, yhat, z, dz * valid, valid, ns

    # ---- B1: the two reductions of the BatchNorm backward ----
    S1 = torch.zeros(N_FILT, dtype=torch.float64)
    S2 = torch.zeros(N_FILT, dtype=torch.float64)
    for b0 in range(0, B, TILE_S):
        for c in range(N_CH):
            _, _, _, yhat, _, dz, valid, _ = tile_quantities(b0, c)
            S1 += dz.double().sum(0)
            S2 += (dz * yhat * valid).double().sum(0)
    m1, m2 = (S1 / n).float(), (S2 / n).float()
    # folded constants: dy = gr*(dz - m1 - yhat*m2) = A*dz + B*y + C   (conv_temporal_bwd v2)
    gr = P["g1"] * rstd
    cA, cB, cC = gr, -gr * m2 * rstd, gr * (m2 * rstd * mean - m1)
    # ---- B2: channel-major ----
    g = {"dwt": torch.zeros(N_FILT, K_TEMP), "dbt": torch.zeros(N_FILT), "dws": torch.zeros(N_FILT, N_FILT, N_CH),
         "dx3": torch.zeros(B, N_CH, N_T), "dg1": S2.float(), "db1": S1.float()}
    wt51 = tf32_rn(P["wt"].reshape(N_FILT, K_TEMP) / K_POOL)
    for c in range(N_CH):
        dws_acc = torch.zeros(48, 48)          # lives in TMEM for the whole sample loop of this CTA
        dwt_acc = torch.zeros(48, 32)
        for b0 in range(0, B, TILE_S):
            dyt, A, y, yhat, z, dz, valid, ns = tile_quantities(b0, c)
            dy = tf32_rn((cA[None, :] * dz + cB[None, :] * y + cC[None, :]) * valid)         # [128, 40]
            a1 = tf32_rn(elu(z)) * valid
            # dWs[j, k] += sum_rows dY2[row, j] a1[row, k]: both operands MN-major views of row-major tiles, K = 128 rows
            a1p = torch.zeros(128, 48)
            a1p[:, :N_FILT] = a1
            dws_acc += mma(dyt.T.contiguous(), a1p.T.contiguous())
            # dWt[k, t] += sum_rows dy[row, k] ps[row, t]   (the 1/51 is applied once at the end)
            dyp = torch.zeros(128, 48)
            dyp[:, :N_FILT] = dy
            dwt_acc += mma(dyp.T.contiguous(), tf32_rn(A).T.contiguous())
            g["dbt"] += dy.sum(0)
            # dps[(s,p), t] = sum_k dy[(s,p), k] wt[k, t]/51: rows (s,p), N = 25 -> 32, K = 40 -> 48
            wtT = torch.zeros(32, 48)
            wtT[:K_TEMP, :N_FILT] = wt51.T
            dps_tile = mma(dyp, wtT)[:, :K_TEMP]                                              # [128, 25]
            for s in range(ns):
                dps = torch.zeros(N_PSUM)
                for p in range(N_POOL):
                    dps[5 * p:5 * p + K_TEMP] += dps_tile[s * N_POOL + p]
                # dx[t] = D[min(t,199)+1] - D[max(t-50,0)], D = exclusive prefix sums of dps
                D = torch.zeros(N_PSUM + 1)
                D[1:] = torch.cumsum(dps, 0)
                t = torch.arange(N_T)
                g["dx3"][b0 + s, c] = D[torch.clamp(t, max=N_PSUM - 1) + 1] - D[torch.clamp(t - (K_POOL - 1), min=0)]
        g["dws"][:, :, c] = dws_acc[:N_FILT, :N_FILT]
        g["dwt"] += dwt_acc[:N_FILT, :K_TEMP] / K_POOL
    return g


# ------------------------------------------------------------------------------------------------ reference (plain torch ops)
def reference(x3, P, dy2):
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    x = x3.clone().requires_grad_(True)
    c = F.conv2d(x.unsqueeze(1), leaves["wt"], leaves["bt"])
    c = F.avg_pool2d(c, (1, K_POOL), (1, S_POOL))                          # (B,40,63,36)
    m = c.mean(dim=(0, 2, 3))
    v = c.var(dim=(0, 2, 3), unbiased=False)
    a = F.elu((c - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + EPS) * leaves["g1"][None, :, None, None]
              + leaves["b1"][None, :, None, None])
    y2 = F.conv2d(a, leaves["ws"], leaves["bs"]).squeeze(2).permute(0, 2, 1)          # (B,36,40) rows (b,p)
    (y2 * dy2).sum().backward()
    return y2.detach(), m.detach(), v.detach(), {"dwt": leaves["wt"].grad.reshape(N_FILT, K_TEMP), "dbt": leaves["bt"].grad,
                                                 "dws": leaves["ws"].grad.reshape(N_FILT, N_FILT, N_CH), "dx3": x.grad,
                                                 "dg1": leaves["g1"].grad, "db1": leaves["b1"].grad}


def main(B=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    x3 = torch.randn(B, N_CH, N_T, generator=g)
    x3 = (x3 - x3.mean(-1, keepdim=True)) / x3.std(-1, keepdim=True)               # LayerNorm-like rows, as the encoder emits
    P = {"wt": (torch.rand(N_FILT, 1, 1, K_TEMP, generator=g) * 2 - 1) * 0.2, "bt": 0.05 * torch.randn(N_FILT, generator=g),
         "g1": 1 + 0.1 * torch.randn(N_FILT, generator=g), "b1": 0.05 * torch.randn(N_FILT, generator=g),
         "ws": (torch.rand(N_FILT, N_FILT, N_CH, 1, generator=g) * 2 - 1) * 0.02, "bs": 0.05 * torch.randn(N_FILT, generator=g)}
    dy2 = torch.randn(B, N_POOL, N_FILT, generator=g) * 0.1
    y2_ref, m_ref, v_ref, g_ref = reference(x3, P, dy2)
    y2, mean, var, rstd = forward_plan(x3, P)
    grads = backward_plan(x3, P, dy2, mean, rstd)

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

    res = {"bn1 mean": (mean - m_ref).abs().max().item(), "bn1 var": rel(var, v_ref), "y2": rel(y2, y2_ref)}
    for k in ("dwt", "dws", "dx3", "dg1", "db1"):
        res[k] = rel(grads[k], g_ref[k])
    res["dbt (analytically 0)"] = grads["dbt"].abs().max().item()
    ok = True
    tol = {"bn1 mean": 1e-5, "bn1 var": 1e-5, "y2": 1e-3, "dwt": 3e-3, "dws": 3e-3, "dx3": 3e-3, "dg1": 3e-3, "db1": 3e-3,
           "dbt (analytically 0)": 1e-2}
    for k, v in res.items():
        flag = v <= tol[k]
        ok &= flag
        print(f"{k:22s} {v:.3e}  {'ok' if flag else 'MISMATCH'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
