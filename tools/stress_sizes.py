"""Large-shape smoke run (index arithmetic beyond 2^31 bytes, ragged lengths, long captions): forward + loss + backward must stay finite."""
import sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__
__graft_entry__.build()
from icd_b200 import synthetic
from icd_b200.losses import attention_caption_loss
from icd_b200.vocabulary import synthetic_vocab
import icd_b200.models.attention as my_att
dev = torch.device("cuda", 0)
for (B, MAXLEN, prec) in [(1024, 52, "bf16"), (768, 40, "fp32x3"), (1536, 30, "bf16")]:
    V = 9490
    p = my_att.AttentionDecoderParams(); p.vocab = synthetic_vocab(V)
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(dev, p); dec = dec.to(dev); dec.precision = prec; dec.train()
    enc = synthetic.features(B, seed=1).to(dev)
    caps, lens = synthetic.captions(B, V, max_len=MAXLEN, seed=1, lengths='ragged')
    caps = caps.to(dev)
    preds, cs, dl, alphas = dec(enc, caps, lens)
    loss = attention_caption_loss(preds, cs, dl, alphas)
    loss.backward()
    torch.cuda.synchronize()
    gn = sum(float(p_.grad.double().norm()) for p_ in dec.parameters() if p_.grad is not None)
    ok = all(torch.isfinite(p_.grad).all().item() for p_ in dec.parameters() if p_.grad is not None)
    print(B, MAXLEN, prec, "loss", float(loss), "finite", ok and torch.isfinite(preds).all().item(), "gradnorm", gn, "mem GB", torch.cuda.max_memory_allocated() / 2**30, flush=True)
    del dec, enc, preds, alphas, loss
    torch.cuda.empty_cache()
