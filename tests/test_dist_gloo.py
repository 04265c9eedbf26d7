"""CPU tier, world_size 2 over gloo: the data-parallel gradient path of icd_b200.parallel.

Each rank takes half of an equal-length batch, computes the decoder gradients (oracle on CPU stands in for the CUDA
decoder — the collective plumbing is device-agnostic), flattens them with FlatParamBuffer and all-reduces.  The
summed, 1/world-scaled gradient must equal the single-process gradient of the full batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _worker(rank, world, port, ret):
    sys.path.insert(0, H.ROOT)
    sys.path.insert(0, os.path.join(H.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        import icd_b200.models.attention as my_att
        from icd_b200.parallel import FlatParamBuffer, all_reduce_gradients
        from icd_b200.vocabulary import synthetic_vocab
        from oracle import decoders as O
        case = dict(H.ATT_CASES["att_small_dropout"], train=False, dropout=0.0, fine_tune_embedding=True)
        dec = H.build_attention_module(case, my_att.AttentionDecoder, my_att.AttentionDecoderParams,
                                       synthetic_vocab(case["V"]))
        buf = FlatParamBuffer(dec)
        enc, caps, lens = H.att_inputs(case)
        B = case["B"]
        sh = B // world
        sl = slice(rank * sh, (rank + 1) * sh)

        def grads(enc_, caps_, lens_):
            w = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
            p, _, dl, a = O.attention_decoder_forward(w, enc_, caps_, lens_)
            O.attention_loss(p, caps_, dl, a).backward()
            return w
        w = grads(enc[sl], caps[sl], lens[sl])
        for (k, p) in dec.named_parameters():
            p.grad = w[k].grad
        buf.gather_grads()
        n = all_reduce_gradients(buf)
        assert n == world
        got = buf.flat_grad / n
        full = grads(enc, caps, lens)
        want = torch.cat([full[k].grad.reshape(-1) for k, _ in dec.named_parameters()])
        err = float((got - want).norm() / want.norm())
        ret[rank] = err
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-5, "rank %d: DP gradient differs from full batch by %.3e" % (r, ret[r])
