"""Diagnostic: pure-write, pure-read and copy bandwidth of this GPU (is a store-only kernel such as the vocabulary layer's
466 MB logit write bound by a write-bandwidth ceiling below the copy figure?)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 28                                   # 1 GiB of fp32
x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
def t(f, iters=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
gb = n * 4 / 1e9
print("fill   (write only): %7.1f GB/s" % (gb / t(lambda: x.fill_(1.0))))
print("memset (write only): %7.1f GB/s" % (gb / t(lambda: x.zero_())))
print("sum    (read only) : %7.1f GB/s" % (gb / t(lambda: x.sum())))
print("copy   (read+write): %7.1f GB/s" % (2 * gb / t(lambda: y.copy_(x))))
h = x.view(torch.bfloat16)[:n]
print("fp32->bf16 convert  : %7.1f GB/s (read 4 B + write 2 B per element)" % (1.5 * gb / t(lambda: torch.empty(n, device=dev, dtype=torch.bfloat16).copy_(x))))
