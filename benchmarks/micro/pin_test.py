import torch, os, time
n = 512 << 20
path = '/dev/shm/ttl_pin_test'
open(path, 'wb').truncate(n)
buf = torch.from_file(path, shared=True, size=n, dtype=torch.uint8)
t0 = time.perf_counter()
rc = torch.cuda.cudart().cudaHostRegister(buf.data_ptr(), n, 1)
print('register rc', rc, 'took %.1f ms' % (1e3 * (time.perf_counter() - t0)), 'is_pinned', buf.is_pinned())
src = torch.empty((n,), dtype=torch.uint8, device='cuda')
pin = torch.empty((n,), dtype=torch.uint8).pin_memory()
for name, dst in (('registered shm', buf), ('torch pinned', pin)):
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(name, it, '%.2f ms device, %.2f ms host call -> %.1f GB/s' % (e0.elapsed_time(e1), 1e3 * t_host, n / e0.elapsed_time(e1) / 1e6))
os.unlink(path)
