import torch, json
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
n = (1 << 21) * 640
buf = torch.empty(n, dtype=torch.float32, device="cuda")
ibuf = buf.view(torch.int32)
out = {}
out["zero_TBps"] = n * 4 / timed(lambda: buf.zero_()) / 1e9
out["fill_const_TBps"] = n * 4 / timed(lambda: buf.fill_(1.2345)) / 1e9
out["arange_TBps"] = n * 4 / timed(lambda: torch.arange(n, out=ibuf)) / 1e9
src = torch.randn(n // 2, device="cuda")
dst = buf[: n // 2]
out["copy_random_rw_TBps"] = 2 * (n // 2) * 4 / timed(lambda: dst.copy_(src)) / 1e9
print(json.dumps(out))
