import torch, time
n_h2d, n_d2h = 133_000_000, 200_000_000
h_in = torch.empty(n_h2d, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n_h2d, dtype=torch.uint8, device='cuda')
h_out = torch.empty(n_d2h, dtype=torch.uint8).pin_memory(); d_out = torch.empty(n_d2h, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps*1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a=t(h2d); b=t(d2h); c=t(both)
print(f"H2D {n_h2d/a/1e6:.1f} GB/s ({a:.2f} ms)  D2H {n_d2h/b/1e6:.1f} GB/s ({b:.2f} ms)  both {c:.2f} ms")
