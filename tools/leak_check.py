"""Does a train step leave cyclic garbage that pins device memory when the cyclic GC is off?"""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
exec(open(os.path.join(ROOT, "tools", "host_profile.py")).read().split("for _ in range(3):")[0])
for _ in range(3):
    step()
torch.cuda.synchronize()
gc.collect()
gc.disable()
for i in range(8):
    step()
    torch.cuda.synchronize()
    st = torch.cuda.memory_stats()
    print("step %d allocated %.2f GB reserved %.2f GB device_allocs %d gc_objects %d" % (
        i, torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30, st["num_device_alloc"], len(gc.get_objects())))
n = gc.collect()
torch.cuda.synchronize()
print("gc.collect() freed %d objects -> allocated %.2f GB" % (n, torch.cuda.memory_allocated() / 2**30))
