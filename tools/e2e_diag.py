"""Diagnostic: which part of bench.py's end-to-end loop costs time over the device-resident loop?
Runs the same train step under a series of loop variants and prints ms/step for each (CUDA events)."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__
__graft_entry__.build()
from icd_b200 import ops, synthetic
from icd_b200.losses import attention_caption_loss
from icd_b200.parallel import DataParallelClipAdam
from icd_b200.vocabulary import synthetic_vocab
import icd_b200.models.attention as my_att

dev = torch.device("cuda", 0)
B, V, MAXLEN = 512, 9490, 25
p = my_att.AttentionDecoderParams(); p.vocab = synthetic_vocab(V)
torch.manual_seed(0)
dec = my_att.AttentionDecoder(dev, p); dec.fine_tune_embeddings(False); dec = dec.to(dev); dec.precision = "bf16"; dec.train()
opt = DataParallelClipAdam(dec, lr=1e-4, grad_clip=5.0)
enc_h = synthetic.features(B, seed=1234).pin_memory()
caps_h, lens = synthetic.captions(B, V, max_len=MAXLEN, seed=1234); caps_h = caps_h.pin_memory()
enc_d = enc_h.to(dev); caps_d = caps_h.to(dev)
enc16_d = enc_d.to(torch.bfloat16)

def train_step(enc, caps):
    preds, cs, dl, alphas = dec(enc, caps, lens)
    loss = attention_caption_loss(preds, cs, dl, alphas, alpha_c=1.0)
    opt.zero_grad(); loss.backward(); opt.step()
    return loss

def timed(name, body, n=20):
    for i in range(4): body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): body(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1)/n:8.3f} ms/step", flush=True)

for _ in range(6): train_step(enc_d, caps_d)
torch.cuda.synchronize(); gc.collect(); gc.disable()

ops.prof_enable(True)
timed("resident fp32 input, prof on", lambda i: train_step(enc_d, caps_d))
ops.prof_collect(); ops.prof_enable(False)
timed("resident fp32 input, prof off", lambda i: train_step(enc_d, caps_d))
if os.environ.get("ICD_DIAG_SHORT"):
    sys.exit(0)
timed("resident bf16 input", lambda i: train_step(enc16_d, caps_d))
bufs = [(enc16_d.clone(), caps_d.clone()) for _ in range(2)]
timed("alternating bf16 buffers", lambda i: train_step(*bufs[i % 2]))

loss_host = torch.zeros(2).pin_memory()
evs = [torch.cuda.Event() for _ in range(2)]
def with_d2h(i):
    l = train_step(*bufs[i % 2])
    loss_host[i % 2:i % 2 + 1].copy_(l.detach().reshape(1), non_blocking=True)
    evs[i % 2].record()
timed("+ async loss D2H, no host sync", with_d2h)
state = {"n": 0}
def with_sync(i):
    l = train_step(*bufs[i % 2])
    if state["n"] >= 1:
        evs[(i - 1) % 2].synchronize()
    loss_host[i % 2:i % 2 + 1].copy_(l.detach().reshape(1), non_blocking=True)
    evs[i % 2].record(); state["n"] += 1
timed("+ host sync on previous loss (pipelined)", with_sync)
def with_sync2(i):
    l = train_step(*bufs[i % 2])
    loss_host[i % 2:i % 2 + 1].copy_(l.detach().reshape(1), non_blocking=True)
    evs[i % 2].record(); state["n"] += 1
    if state["n"] >= 3:
        evs[(i - 1) % 2].synchronize()
timed("+ host sync on previous loss AFTER issuing copy", with_sync2)
def with_item(i):
    l = train_step(*bufs[i % 2]); l.item()
timed("loss.item() every step (reference style)", with_item)
cs = torch.cuda.Stream()
ready = [torch.cuda.Event() for _ in range(2)]; done = [torch.cuda.Event() for _ in range(2)]
for d_ in done: d_.record()
enc_h16 = enc_h.to(torch.bfloat16).pin_memory()
def prefetch(i):
    with torch.cuda.stream(cs):
        cs.wait_event(done[i % 2])
        bufs[i % 2][0].copy_(enc_h16, non_blocking=True); bufs[i % 2][1].copy_(caps_h, non_blocking=True)
        ready[i % 2].record(cs)
prefetch(0)
def full(i):
    prefetch(i + 1)
    torch.cuda.current_stream().wait_event(ready[i % 2])
    l = train_step(*bufs[i % 2]); done[i % 2].record()
    loss_host[i % 2:i % 2 + 1].copy_(l.detach().reshape(1), non_blocking=True)
timed("H2D prefetch on side stream, no host sync", full)
timed("resident fp32 input again", lambda i: train_step(enc_d, caps_d))
