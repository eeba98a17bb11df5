"""Development tool: host-to-device copy time of one scan (3.84 MB pinned) on this box: one copy, four chunks on one
stream, four chunks over two streams, and the asymptotic rate of a 64 MB copy."""
import torch
n = 240000 * 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
big = torch.empty(16 * 1024 * 1024, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
dbig = torch.empty_like(big, device="cuda")
s0 = torch.cuda.Stream(); s1 = torch.cuda.Stream(); main = torch.cuda.current_stream()
def timed(fn, reps=20):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        fn()
        e1.record(main)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
def one():
    d.copy_(h, non_blocking=True)
def four_one_stream():
    q = n // 4
    for c in range(4):
        d[c * q:(c + 1) * q].copy_(h[c * q:(c + 1) * q], non_blocking=True)
def four_two_streams():
    q = n // 4
    ev = torch.cuda.Event(); ev.record(main)
    for c in range(4):
        s = s0 if c % 2 == 0 else s1
        s.wait_event(ev)
        with torch.cuda.stream(s):
            d[c * q:(c + 1) * q].copy_(h[c * q:(c + 1) * q], non_blocking=True)
    main.wait_stream(s0); main.wait_stream(s1)
def bigc():
    dbig.copy_(big, non_blocking=True)
for name, fn, nbytes in (("one 3.84 MB copy", one, n * 4), ("4 chunks, one stream", four_one_stream, n * 4),
                         ("4 chunks, two streams", four_two_streams, n * 4), ("64 MB copy", bigc, big.numel() * 4)):
    med, mn = timed(fn)
    print(f"{name}: median {med:.1f} us (min {mn:.1f}) -> {nbytes / med / 1e3:.1f} GB/s")
