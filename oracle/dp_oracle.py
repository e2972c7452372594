"""CPU restatement of the data-parallel protocol used by eeg_image_decode_b200/train.py::StepEngine (TEST INFRASTRUCTURE,
same rules as atms_oracle.py).  Each function is what ONE rank does, with real torch.distributed collectives, so the
world_size-2 gloo tests can check "W ranks x B_local == one process at batch W*B_local"  against
the single-process oracle, i.e. against the reference's world_size>1, local_loss=False semantics
(models/loss.py:59-73, 113-120).
"""
import torch
import torch.distributed as dist


def row_block_infonce(e_local, img_local, txt_local, logit_scale, alpha=0.99):
    """mirrors loss.py::_InfoNCE.run + csrc/loss.cu: gather targets, local B x N logits, row LSE, per-column (max, sum)
    partials exchanged with one all-gather, loss share, G, dE, d(logit_scale) share."""
    W, rank = dist.get_world_size(), dist.get_rank()
    B = e_local.shape[0]
    N = W * B

    def gather(t):
        out = [torch.empty_like(t) for _ in range(W)]
        dist.all_gather(out, t.contiguous())
        return torch.cat(out, 0)

    t_cat = torch.cat([gather(img_local), gather(txt_local)], 0)            # [2N, D]
    logits = logit_scale * e_local @ t_cat.T                                # [B, 2N]
    w = torch.tensor([alpha, 1 - alpha])
    loss_share = torch.zeros(3)
    g = torch.zeros_like(logits)
    # column partials (running max, sum-exp) -> all-gather -> merged column LSE
    cmax = logits.max(dim=0).values
    csum = torch.exp(logits - cmax).sum(dim=0)
    parts = [torch.empty(2, 2 * N) for _ in range(W)]
    dist.all_gather(parts, torch.stack([cmax, csum]))
    parts = torch.stack(parts)                                              # [W, 2, 2N]
    M = parts[:, 0].max(dim=0).values
    S = (parts[:, 1] * torch.exp(parts[:, 0] - M)).sum(dim=0)
    col_lse = M + torch.log(S)
    for t in range(2):
        blk = logits[:, t * N:(t + 1) * N]
        row_lse = torch.logsumexp(blk, dim=1)
        idx = torch.arange(B) + rank * B
        diag = blk[torch.arange(B), idx]
        share = ((row_lse - diag) + (col_lse[t * N + idx] - diag)).sum() / (2 * N)
        loss_share[1 + t] = share
        gt = torch.exp(blk - row_lse[:, None]) + torch.exp(blk - col_lse[None, t * N:(t + 1) * N])
        gt[torch.arange(B), idx] -= 2.0
        g[:, t * N:(t + 1) * N] = gt * (w[t] / (2 * N))
    loss_share[0] = alpha * loss_share[1] + (1 - alpha) * loss_share[2]
    d_e = logit_scale * g @ t_cat
    d_s = (g * logits).sum() / logit_scale
    return loss_share, d_e, d_s


def sync_batch_stats(y_local):
    """SyncBN statistics exchange of StepEngine.step: all-reduce (sum, sum of squares) per channel (dim 1)"""
    dims = [d for d in range(y_local.dim()) if d != 1]
    sums = torch.stack([y_local.double().sum(dim=dims), (y_local.double() ** 2).sum(dim=dims)])
    dist.all_reduce(sums)
    count = y_local.numel() // y_local.shape[1] * dist.get_world_size()
    mean = sums[0] / count
    var = sums[1] / count - mean * mean
    return mean.float(), var.float(), count


def reconstruction_share(e_local, img_local, logit_scale, alpha=0.90):
    """data-parallel form of the reconstruction-training loss (Generation/ATMS_reconstruction.py:227-228) as
    StepEngine.loss_and_grad(variant="reconstruction") composes it: (1-alpha)*10 x the row-block ClipLoss(img) share plus
    alpha*10 x this rank's rows of the global-batch MSE mean.  Returns (loss share, d loss / d e_local)."""
    W = dist.get_world_size()
    n_total = W * e_local.shape[0]
    # single-target row-block InfoNCE == row_block_infonce with all the weight on the image target
    share, d_e, _ = row_block_infonce(e_local, img_local, img_local, logit_scale, alpha=1.0)
    w_clip, w_mse = (1.0 - alpha) * 10.0, alpha * 10.0
    diff = e_local - img_local
    mse_share = (diff * diff).sum() / (n_total * e_local.shape[1])
    return w_clip * share[1] + w_mse * mse_share, w_clip * d_e + w_mse * 2.0 * diff / (n_total * e_local.shape[1])
