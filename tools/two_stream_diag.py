"""Feasibility check: do two half-batches on two streams overlap the HBM-bound attention kernels of one half with the
latency-bound contractions of the other?  Times forward + loss + backward (no optimizer) for
 (a) one 512-caption batch, (b) two 256-caption halves back to back on one stream, (c) the halves on two streams."""
import os, sys, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__
__graft_entry__.build()
from icd_b200 import synthetic
from icd_b200.losses import attention_caption_loss
from icd_b200.vocabulary import synthetic_vocab
import icd_b200.models.attention as my_att

dev = torch.device("cuda", 0)
B, V, MAXLEN = 512, 9490, 25
p = my_att.AttentionDecoderParams(); p.vocab = synthetic_vocab(V)
torch.manual_seed(0)
dec = my_att.AttentionDecoder(dev, p); dec.fine_tune_embeddings(False); dec = dec.to(dev); dec.precision = "bf16"; dec.train()
enc = synthetic.features(B, seed=1234).to(dev)
caps, lens = synthetic.captions(B, V, max_len=MAXLEN, seed=1234); caps = caps.to(dev)
lens_t = torch.as_tensor(lens)
# interleave so both halves have the same length distribution
idx = [torch.arange(0, B, 2), torch.arange(1, B, 2)]
halves = [(enc[i.to(dev)].contiguous(), caps[i.to(dev)].contiguous(), [lens[j] for j in i.tolist()] if not torch.is_tensor(lens) else lens[i]) for i in idx]

def step(e, c, l):
    preds, cs, dl, alphas = dec(e, c, l)
    loss = attention_caption_loss(preds, cs, dl, alphas, alpha_c=1.0)
    loss.backward()
    return loss

def timed(name, body, n=10):
    for _ in range(4): body()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): body()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:55s} {e0.elapsed_time(e1)/n:8.3f} ms", flush=True)

gc.disable()
timed("(a) one batch of 512", lambda: step(enc, caps, lens))
timed("(b) two halves of 256, one stream", lambda: [step(*h) for h in halves])
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): step(*halves[0])
    with torch.cuda.stream(s2): step(*halves[1])
    cur.wait_stream(s1); cur.wait_stream(s2)
timed("(c) two halves of 256, two streams", two)
import threading
def two_threads():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    def w(s, h):
        with torch.cuda.stream(s): step(*h)
    ts = [threading.Thread(target=w, args=(s1, halves[0])), threading.Thread(target=w, args=(s2, halves[1]))]
    for t in ts: t.start()
    for t in ts: t.join()
    cur.wait_stream(s1); cur.wait_stream(s2)
timed("(d) two halves, two streams, two host threads", two_threads)
