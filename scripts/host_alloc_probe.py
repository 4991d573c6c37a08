import time, mmap, numpy as np, torch
n = 514 * 514 * 514 * 3
def t(f, name):
    t0 = time.perf_counter(); r = f(); dt = time.perf_counter() - t0; print("%-32s %.3f s" % (name, dt)); return r
def plain():
    a = np.empty(n, np.float32); a[::1024] = 1; return a
def huge():
    m = mmap.mmap(-1, n * 4, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        m.madvise(mmap.MADV_HUGEPAGE)
    except Exception as e:
        print("madvise:", e)
    a = np.frombuffer(m, np.float32); a[::1024] = 1; return a
def pinned():
    return torch.empty(n, dtype=torch.float32, pin_memory=True)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
torch.cuda.init()
for k in range(2):
    a = t(plain, "np.empty + touch"); del a
    a = t(huge, "mmap+MADV_HUGEPAGE + touch"); del a
    a = t(pinned, "torch pinned alloc"); 
    d = torch.empty(n, dtype=torch.float32, device="cuda"); torch.cuda.synchronize()
    t(lambda: (a.copy_(d), torch.cuda.synchronize()), "D2H into pinned")
    b = np.empty(n, np.float32)
    t(lambda: (torch.from_numpy(b).copy_(d), torch.cuda.synchronize()), "D2H into fresh pageable")
    t(lambda: (torch.from_numpy(b).copy_(d), torch.cuda.synchronize()), "D2H into touched pageable")
    del a, b, d
