"""Host-side profile of the benchmark train step (where does the launching thread spend its time?)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from icd_b200 import synthetic  # noqa: E402
from icd_b200.losses import attention_caption_loss  # noqa: E402
from icd_b200.parallel import DataParallelClipAdam  # noqa: E402
from icd_b200.vocabulary import synthetic_vocab  # noqa: E402
import icd_b200.models.attention as my_att  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
p = my_att.AttentionDecoderParams()
p.vocab = synthetic_vocab(bench.V)
torch.manual_seed(0)
dec = my_att.AttentionDecoder(dev, p)
dec.fine_tune_embeddings(False)
dec = dec.to(dev)
dec.precision = prec
dec.train()
opt = DataParallelClipAdam(dec)
enc = synthetic.features(B).to(dev)
caps, lens = synthetic.captions(B, bench.V, max_len=bench.MAXLEN)
caps = caps.to(dev)


def step():
    preds, cs, dl, alphas = dec(enc, caps, lens)
    loss = attention_caption_loss(preds, cs, dl, alphas)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
pr.disable()
t_all = time.perf_counter() - t0
print("host-side time per step %.2f ms, wall per step %.2f ms" % (t_host / 5 * 1e3, t_all / 5 * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
