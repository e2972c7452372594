"""world_size-2 gloo (CPU) tests of the data-parallel protocol: row-block InfoNCE with the
column-statistics exchange and SyncBN statistics, each checked against the single-process oracle at the global batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import recipe
    from oracle import atms_oracle as O
    from oracle import dp_oracle as DP
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        B, N = 6, 6 * world
        g = torch.Generator().manual_seed(5)
        E = torch.randn(N, 1024, generator=g)
        img = recipe.make_targets(N, seed=3, tag="img")
        txt = recipe.make_targets(N, seed=3, tag="txt")
        s = torch.tensor(2.659)
        sl = slice(rank * B, (rank + 1) * B)
        share, d_e, d_s = DP.row_block_infonce(E[sl], img[sl], txt[sl], s)
        tot = share.clone()
        dist.all_reduce(tot)
        ds_tot = d_s.clone()
        dist.all_reduce(ds_tot)
        Eg = E.clone().requires_grad_(True)
        sg = s.clone().requires_grad_(True)
        ref = O.contrastive_loss(Eg, img, txt, sg)            # == clip_loss_global over the gathered chunks
        ref.backward()
        ref2 = 0.99 * O.clip_loss_global([E[:B], E[B:]], [img[:B], img[B:]], s) + 0.01 * O.clip_loss_global([E[:B], E[B:]], [txt[:B], txt[B:]], s)
        res = {
            "loss": abs(tot[0].item() - ref.item()) / abs(ref.item()),
            "loss_vs_gathered": abs(ref2.item() - ref.item()),
            "dE": ((d_e - Eg.grad[sl]).norm() / Eg.grad[sl].norm()).item(),
            "ds": abs(ds_tot.item() - sg.grad.item()) / abs(sg.grad.item()),
        }
        # reconstruction-training variant: MSE is a mean over the GLOBAL batch, every rank adds its rows' share
        r_share, r_de = DP.reconstruction_share(E[sl], img[sl], s, alpha=0.90)
        r_tot = r_share.clone()
        dist.all_reduce(r_tot)
        Er = E.clone().requires_grad_(True)
        r_ref = O.reconstruction_loss(Er, img, s, 0.90)
        r_ref.backward()
        res["recon_loss"] = abs(r_tot.item() - r_ref.item()) / abs(r_ref.item())
        res["recon_dE"] = ((r_de - Er.grad[sl]).norm() / Er.grad[sl].norm()).item()
        # SyncBN statistics == statistics of the global batch
        y = torch.randn(N, 40, 63, 36, generator=g)
        mean, var, count = DP.sync_batch_stats(y[sl])
        res["bn_mean"] = (mean - y.mean(dim=(0, 2, 3))).abs().max().item()
        res["bn_var"] = (var - y.var(dim=(0, 2, 3), unbiased=False)).abs().max().item()
        res["bn_count"] = count - N * 63 * 36
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_two_rank_protocol_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = q.get()
    assert res["loss"] < 1e-5 and res["loss_vs_gathered"] < 1e-6, res
    assert res["dE"] < 1e-4 and res["ds"] < 1e-4, res
    assert res["recon_loss"] < 1e-5 and res["recon_dE"] < 1e-4, res
    assert res["bn_mean"] < 1e-5 and res["bn_var"] < 1e-4 and res["bn_count"] == 0, res


def test_graphed_step_and_engine_are_importable_without_cuda():
    from eeg_image_decode_b200.train import GraphedTrainStep, StepEngine, fused_adamw_step  # noqa: F401
