import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
print("ALLOC_CONF", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_ALLOC_CONF"))
print(torch.cuda.get_allocator_backend(), torch.cuda.mem_get_info())
x = []
torch.cuda.synchronize()
for sz in (1 << 20, 100 << 20, 466 << 20):
    t0 = time.perf_counter(); a = torch.empty(sz, dtype=torch.uint8, device="cuda"); t1 = time.perf_counter()
    del a
    t2 = time.perf_counter(); a = torch.empty(sz, dtype=torch.uint8, device="cuda"); t3 = time.perf_counter()
    print(sz >> 20, "MB first %.3f ms, cached %.3f ms" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3))
    del a
exec(open(os.path.join(ROOT, "tools", "host_profile.py")).read().split("for _ in range(3):")[0])
st0 = torch.cuda.memory_stats()
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    st = torch.cuda.memory_stats()
    print("step %d host %.1f ms total %.1f ms  device_alloc %d device_free %d retries %d reserved %.1f GB allocated-peak %.1f GB" % (
        i, (t1 - t0) * 1e3, (t2 - t0) * 1e3, st["num_device_alloc"], st["num_device_free"], st["num_alloc_retries"],
        st["reserved_bytes.all.current"] / 2**30, st["allocated_bytes.all.peak"] / 2**30))
