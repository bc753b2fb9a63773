import torch
dev = torch.device("cuda", 0)
n = 1 << 29  # 512 Mi elements of fp16 = 1 GiB
a = torch.empty(n, dtype=torch.float16, device=dev); b = torch.empty(n, dtype=torch.float16, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e-3
GB = n * 2 / 1e9
print("fill  (write only) %.0f GB/s" % (GB / t(lambda: a.fill_(1.0))))
print("copy  (r+w)        %.0f GB/s" % (2 * GB / t(lambda: b.copy_(a))))
print("sum   (read only)  %.0f GB/s" % (GB / t(lambda: a.view(torch.int16).max())))
# 1:4 read:write  (expand-like): write 4 outputs from one read
c = torch.empty(4, n // 4, dtype=torch.float16, device=dev); d = a[: n // 4]
print("1r:4w broadcast    %.0f GB/s" % ((GB / 4 + GB) / t(lambda: c.copy_(d.unsqueeze(0).expand(4, -1)))))
print("memset             %.0f GB/s" % (GB / t(lambda: a.zero_())))
