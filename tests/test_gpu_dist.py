"""GPU tier, 2 ranks over NCCL (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`):
the CUDA decoders + fused loss + NCCL gradient all-reduce + fused 1/world-scale / clamp / Adam kernel — the whole
data-parallel train step of icd_b200.parallel.

Checked per decoder (attention: bf16 tier with bf16_grad_only, train-mode dropout from a shared keep-mask; baseline: bf16
tier):
  * the all-reduced, 1/world-scaled flat gradient of 2 ranks x B/2 rows equals the single-process gradient of the B rows
    (equal-length captions: every rank holds the same number of packed tokens, so mean-of-means == global mean);
  * after the optimiser step the parameters are BIT-identical on both ranks;
  * they match the single-process step (Adam's first step is lr * g / (|g| + eps): elements whose gradient is ~0 may
    differ by up to 2 lr, everything else agrees to 1e-2 lr on average).
"""
import os
import sys

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _attention_setup(dev):
    import icd_b200.models.attention as my_att
    from icd_b200 import synthetic
    from icd_b200.vocabulary import synthetic_vocab
    V, B, L = 1000, 64, 12
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    p.attention_dim, p.decoder_dim, p.embed_size = 128, 128, 64

    def make():
        torch.manual_seed(0)
        dec = my_att.AttentionDecoder(dev, p)
        dec.fine_tune_embeddings(True)
        dec = dec.to(dev)
        dec.precision = "bf16"
        dec.train()
        return dec
    enc = synthetic.features(B, seed=5)
    caps, lens = synthetic.captions(B, V, max_len=L, seed=5)
    keep = H.seeded_keep_mask(L - 1, B, 128, 0.5, seed=8)
    return make, enc, caps, lens, keep


def _attention_step(dec, opt, enc, caps, lens, keep, dev):
    from icd_b200.losses import attention_caption_loss
    dec._dropout_mask_override = keep
    preds, cs, dl, alphas = dec(enc.to(dev), caps.to(dev), lens)
    loss = attention_caption_loss(preds, cs, dl, alphas, bf16_grad_only=True)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


def _baseline_setup(dev):
    import icd_b200.models.baseline as my_base
    from icd_b200 import synthetic
    V, B, L = 1000, 64, 12
    p = my_base.BaselineDecoderParams()
    p.vocab_size, p.embed_size, p.hidden_size = V, 64, 128

    def make():
        torch.manual_seed(0)
        dec = my_base.BaselineDecoder(p).to(dev)
        dec.precision = "bf16"
        return dec
    g = torch.Generator().manual_seed(3)
    img = torch.randn(B, 64, generator=g)
    caps, lens = synthetic.captions(B, V, max_len=L, seed=6)
    return make, img, caps


def _baseline_step(dec, opt, img, caps, dev):
    from icd_b200.losses import baseline_caption_loss
    out = dec(img.to(dev), caps.to(dev))
    loss = baseline_caption_loss(out, caps.to(dev))
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, H.ROOT)
    sys.path.insert(0, os.path.join(H.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from icd_b200.parallel import DataParallelClipAdam
        lr = 1e-3
        res = {}
        for which in ("attention", "baseline"):
            if which == "attention":
                make, x, caps, lens, keep = _attention_setup(dev)
            else:
                make, x, caps = _baseline_setup(dev)
            B = x.shape[0]
            sh = B // world
            sl = slice(rank * sh, (rank + 1) * sh)
            dec = make()
            opt = DataParallelClipAdam(dec, lr=lr, grad_clip=5.0)
            if which == "attention":
                _attention_step(dec, opt, x[sl], caps[sl], lens[sl], keep[:, sl].contiguous(), dev)
            else:
                _baseline_step(dec, opt, x[sl], caps[sl], dev)
            flat, grad = opt.buf.flat.clone(), opt.buf.flat_grad.clone() / world
            both = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(both, flat)
            res[which + "_ranks_bit_identical"] = bool(all(torch.equal(both[0], b) for b in both))
            # single-process step over the whole batch, on this rank's GPU, without any collective
            ref = make()
            ropt = DataParallelClipAdam(ref, lr=lr, grad_clip=5.0, group=False)
            if which == "attention":
                _attention_step(ref, ropt, x, caps, lens, keep, dev)
            else:
                _baseline_step(ref, ropt, x, caps, dev)
            res[which + "_grad_rel_err"] = H.rel_err(grad, ropt.buf.flat_grad)
            dp = (flat - ropt.buf.flat).abs()
            res[which + "_param_max_abs_diff_over_lr"] = float(dp.max()) / lr
            res[which + "_param_mean_abs_diff_over_lr"] = float(dp.mean()) / lr
        ret[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_gpu_train_step_equals_single_process_step():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    print("\n2-GPU data-parallel equivalence:", dict(ret[0]))
    out = os.environ.get("ICD_DP_REPORT")
    if out:
        import json
        with open(out, "w") as f:
            json.dump({str(r): dict(ret[r]) for r in range(world)}, f, indent=1)
    for r in range(world):
        for which in ("attention", "baseline"):
            assert ret[r][which + "_ranks_bit_identical"], which
            assert ret[r][which + "_grad_rel_err"] < 2e-5, (which, ret[r][which + "_grad_rel_err"])
            assert ret[r][which + "_param_max_abs_diff_over_lr"] <= 2.1
            assert ret[r][which + "_param_mean_abs_diff_over_lr"] < 1e-2
