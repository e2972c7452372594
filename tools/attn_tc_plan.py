"""Executable restatement (CPU, torch) of the block-diagonal tcgen05 blocking of the attention core
(csrc/attention_tc.cu forward, verified on B200; backward = round-2 plan): two samples of one
head form one 128-row problem, the off-diagonal 64x64 blocks of S / dP are ignored and those of P / dS are zero, so that
every product is ONE M=128 UMMA group.  Checked against autograd through the reference formulation
(SelfAttention_Family.py:56-75).      python tools/attn_tc_plan.py
"""
import math
import sys

import torch

L, E = 64, 62            # tokens per sample, head dim (padded to 64 in the kernels)
SCALE = 1.0 / math.sqrt(E)


def tf32_rn(x):
    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def mma(a, b):           # a [M,K] . b[N,K]^T, TF32 operands, fp32 accumulate
    return a.double().matmul(b.double().T).float()


def pad64(x):            # [.., 62] -> [.., 64]
    return torch.nn.functional.pad(x, (0, 64 - x.shape[-1]))


def pair_forward(q, k, v, keep, p_drop):
    """q, k, v: [2, 64, 62] (two samples, one head).  keep: [2, 64, 64] {0,1}.  Returns O [2, 64, 62] and P (dropped)."""
    Q, K, V = (tf32_rn(pad64(t).reshape(128, 64)) for t in (q, k, v))
    S = mma(Q, K)                                        # [128, 128]; only the diagonal blocks are read
    P = torch.zeros(128, 128)
    for h in range(2):
        blk = S[64 * h:64 * h + 64, 64 * h:64 * h + 64] * SCALE
        pr = torch.softmax(blk, dim=-1) * keep[h] / (1.0 - p_drop)
        P[64 * h:64 * h + 64, 64 * h:64 * h + 64] = pr
    P = tf32_rn(P)
    O = mma(P, V.T.contiguous())                         # A = P (K-major over the 128 keys), B = V as MN-major operand
    return O.reshape(2, 64, 64)[..., :E], P


def pair_backward(q, k, v, d_o, keep, p_drop):
    """gradients wrt q, k, v for one pair; every product is a 128-row tile"""
    Q, K, V, DO = (tf32_rn(pad64(t).reshape(128, 64)) for t in (q, k, v, d_o))
    S = mma(Q, K)
    dP = mma(DO, V)                                      # dP[i,j] = dO[i,:] . V[j,:]
    P = torch.zeros(128, 128)
    Pd = torch.zeros(128, 128)
    dS = torch.zeros(128, 128)
    for h in range(2):
        sl = slice(64 * h, 64 * h + 64)
        p = torch.softmax(S[sl, sl] * SCALE, dim=-1)
        kf = keep[h] / (1.0 - p_drop)
        dp = dP[sl, sl] * kf                              # gradient wrt the un-dropped probability
        r = (dp * p).sum(-1, keepdim=True)
        P[sl, sl] = p
        Pd[sl, sl] = p * kf
        dS[sl, sl] = p * (dp - r) * SCALE
    Pd, dS = tf32_rn(Pd), tf32_rn(dS)
    dQ = mma(dS, K.T.contiguous())                       # A = dS K-major, B = K MN-major
    dK = mma(dS.T.contiguous(), Q.T.contiguous())        # A = dS^T (= dS tile read MN-major), B = Q MN-major
    dV = mma(Pd.T.contiguous(), DO.T.contiguous())       # A = Pd^T, B = dO MN-major
    f = lambda t: t.reshape(2, 64, 64)[..., :E]
    return f(dQ), f(dK), f(dV)


def reference(q, k, v, d_o, keep, p_drop):
    q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("ble,bse->bls", q, k) * SCALE
    a = torch.softmax(s, dim=-1) * keep / (1.0 - p_drop)
    o = torch.einsum("bls,bsd->bld", a, v)
    (o * d_o).sum().backward()
    return o.detach(), q.grad, k.grad, v.grad


def main(seed=0, p_drop=0.25):
    g = torch.Generator().manual_seed(seed)
    q, k, v, d_o = (torch.randn(2, L, E, generator=g) for _ in range(4))
    keep = (torch.rand(2, L, L, generator=g) >= p_drop).float()
    o_ref, dq_ref, dk_ref, dv_ref = reference(q, k, v, d_o, keep, p_drop)
    o, _ = pair_forward(q, k, v, keep, p_drop)
    dq, dk, dv = pair_backward(q, k, v, d_o, keep, p_drop)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    ok = True
    for name, a, b in (("O", o, o_ref), ("dQ", dq, dq_ref), ("dK", dk, dk_ref), ("dV", dv, dv_ref)):
        e = rel(a, b)
        ok &= e < 3e-3
        print(f"{name:3s} rel err {e:.3e}  {'ok' if e < 3e-3 else 'MISMATCH'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
