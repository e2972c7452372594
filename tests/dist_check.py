"""Data-parallel equivalence check (run under torchrun, one process per GPU):
W ranks x B_local  ==  one process at batch W*B_local  (loss, embeddings, every gradient, BN running stats, updated
parameters).  Both sides run the exact-fp32 verification backend so the comparison is tight (1e-4), then the product
(TF32) backend is checked against the same reference with TF32 tolerances."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import recipe  # noqa: E402
from oracle import atms_oracle as O  # noqa: E402  (the checker: CPU restatement of the reference step)
from eeg_image_decode_b200 import _lib  # noqa: E402
from eeg_image_decode_b200.atms import ATMS  # noqa: E402
from eeg_image_decode_b200.train import StepEngine  # noqa: E402


def make_model(p_drop):
    m = ATMS()
    m.load_state_dict(recipe.make_state_dict())
    m = m.cuda().train()
    m.dropout_p = [p_drop] * 8
    return m


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    Bl = 8
    N = Bl * world
    x = recipe.make_eeg(N, seed=71).cuda()
    sid = torch.full((N,), 8).cuda()
    img = recipe.make_targets(N, seed=71, tag="img").cuda()
    txt = recipe.make_targets(N, seed=71, tag="txt").cuda()
    sl = slice(rank * Bl, (rank + 1) * Bl)
    ok = True
    # the reference semantics of the W-rank step == ONE process at batch W*B_local: the CPU oracle's autograd step
    o_loss, o_grads, o_r = O.train_step(recipe.make_state_dict(), {}, x.cpu(), sid.cpu(), img.cpu(), txt.cpu(), 1)
    for backend, tol_f, tol_g in ((1, 5e-5, 2e-3), (0, 1.5e-3, 4e-2)):
        _lib.set_gemm_backend(backend)
        # --- data parallel
        m = make_model(0.0)
        eng = StepEngine(m, None)
        assert eng.world == world
        loss, feats = eng.step(x[sl], sid[sl], img[sl], txt[sl], use_shared=False)
        total = loss.clone()
        dist.all_reduce(total)
        # --- single process at the global batch (same library, no collectives)
        ref = make_model(0.0)
        e1 = StepEngine(ref, None)
        e1.world, e1.rank = 1, 0
        rloss, rfeats = e1.step(x, sid, img, txt, use_shared=False)
        torch.cuda.synchronize()

        def rel(a, b):
            return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

        errs = {"loss": abs(total[0].item() - rloss[0].item()) / abs(rloss[0].item()), "feats": rel(feats, rfeats[sl])}
        worst_g = 0.0
        for k in ("proj_eeg.0.weight", "proj_eeg.2.weight", "enc_eeg.0.tsconv.0.weight", "enc_eeg.0.tsconv.2.weight",
                  "enc_eeg.0.tsconv.4.weight", "enc_eeg.0.tsconv.5.bias", "enc_eeg.0.projection.0.weight",
                  "encoder.encoder.attn_layers.0.attention.query_projection.weight",
                  "encoder.encoder.attn_layers.0.conv1.weight", "encoder.encoder.attn_layers.0.norm1.weight",
                  "encoder.enc_embedding.value_embedding.weight",
                  "encoder.enc_embedding.subject_embedding.subject_embedding.weight", "logit_scale"):
            worst_g = max(worst_g, rel(m.grad_view(k), ref.grad_view(k)))
        errs["grads"] = worst_g
        bn = max(rel(m.state_dict()[k], ref.state_dict()[k]) for k in
                 ("enc_eeg.0.tsconv.2.running_mean", "enc_eeg.0.tsconv.2.running_var", "enc_eeg.0.tsconv.5.running_var"))
        errs["bn_running"] = bn
        # ... and against the oracle (dp_oracle.py / atms_oracle.py), not only against the library itself
        errs["loss_vs_oracle"] = abs(total[0].item() - o_loss.item()) / abs(o_loss.item())
        errs["feats_vs_oracle"] = rel(feats.cpu(), o_r["out"].detach()[sl])
        errs["grads_vs_oracle"] = max(rel(m.grad_view(k).cpu(), g) for k, g in o_grads.items()
                                      if g is not None and k not in ("enc_eeg.0.tsconv.0.bias", "enc_eeg.0.tsconv.4.bias",
                                                                     "encoder.encoder.attn_layers.0.attention.key_projection.bias"))
        good = errs["loss"] < tol_f * 10 and errs["feats"] < tol_f and errs["grads"] < tol_g and bn < tol_f * 10
        good = good and errs["loss_vs_oracle"] < max(tol_f * 10, 2e-4) and errs["feats_vs_oracle"] < max(tol_f, 2e-5) \
            and errs["grads_vs_oracle"] < tol_g
        # ranks stay in sync after the update
        w = m.state_dict()["proj_eeg.0.weight"].clone()
        w0 = w.clone()
        dist.broadcast(w0, 0)
        sync = torch.equal(w, w0)
        if rank == 0:
            print(f"backend={backend} world={world} errs={errs} ranks_in_sync={sync} -> {'PASS' if good and sync else 'FAIL'}", flush=True)
        ok = ok and good and sync
    _lib.set_gemm_backend(0)
    # dropout on: just has to run and stay finite / in sync
    m = make_model(0.25)
    eng = StepEngine(m, None)
    loss, feats = eng.step(x[sl], sid[sl], img[sl], txt[sl], use_shared=False)
    fin = bool(torch.isfinite(loss).all().item()) and bool(torch.isfinite(m.flat_params).all().item())
    if rank == 0:
        print(f"dropout step finite={fin}", flush=True)
    ok = ok and fin
    # CUDA-graph captured data-parallel step (NCCL collectives inside the graph) vs eager data-parallel, 5 steps
    from eeg_image_decode_b200.train import GraphedTrainStep
    gal = recipe.make_targets(50, seed=71, tag="gal").cuda()
    lab = recipe.make_labels(Bl, 50, seed=71).cuda()
    finals = {}
    for graphed in (False, True):
        mg = make_model(0.0)
        gs = GraphedTrainStep(StepEngine(mg, None), gal, use_shared=False, enabled=graphed)
        ls = []
        for i in range(5):
            l, f, c = gs(x[sl], sid[sl], img[sl], txt[sl], lab)
            t_ = l.clone(); dist.all_reduce(t_); ls.append(t_[0].item())
        finals[graphed] = (ls, mg.flat_params.clone(), gs.graph is not None)
    gerr = max(abs(a - b) / abs(b) for a, b in zip(finals[True][0], finals[False][0]))
    gok = finals[True][2] and not finals[False][2] and gerr < 5e-3
    if rank == 0:
        print(f"graphed DP: captured={finals[True][2]} loss rel diff vs eager={gerr:.2e} -> {'PASS' if gok else 'FAIL'}", flush=True)
    ok = ok and gok
    torch.cuda.synchronize()
    eng.check_collectives()          # the NVLink peer SyncBN exchange (csrc/peer_sum.cu) never gave up on a rank
    if rank == 0:
        print("DIST_CHECK " + ("PASS" if ok else "FAIL"), flush=True)
    # destroy_process_group() with captured collectives alive hung on this stack: drop the graphs, run the interpreter's
    # exit hooks by hand (so that whoever records loaded libraries at exit still does), then leave without destructors
    del gs, mg, finals
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    import atexit
    atexit._run_exitfuncs()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
