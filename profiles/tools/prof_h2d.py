"""Host->device bandwidth of one 1 GiB minibatch: torch pinned memory, cudaHostAlloc write-combined, one / two / four streams.
    python profiles/tools/prof_h2d.py"""
import ctypes, time
import torch
rt = ctypes.CDLL('libcudart.so.12')
N = 1 << 30
dev = torch.device('cuda:0')
dst = torch.empty(N, dtype=torch.uint8, device=dev)
def timed_copy(src_ptr, nstreams, reps=5):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    chunk = N // nstreams
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, s in enumerate(streams):
            s.wait_event(e0)
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + i * chunk), ctypes.c_void_p(src_ptr + i * chunk), ctypes.c_size_t(chunk), 1, ctypes.c_void_p(s.cuda_stream))
            assert rc == 0, rc
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return N / best / 1e6
pin = torch.empty(N, dtype=torch.uint8).pin_memory(); pin.fill_(3)
for ns in (1, 2, 4):
    print('torch pinned, %d stream(s): %.1f GB/s' % (ns, timed_copy(pin.data_ptr(), ns)))
for flags, name in ((0, 'cudaHostAllocDefault'), (4, 'cudaHostAllocWriteCombined')):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), flags); assert rc == 0, rc
    ctypes.memset(p, 5, N)
    for ns in (1, 2):
        print('%s, %d stream(s): %.1f GB/s' % (name, ns, timed_copy(p.value, ns)))
    rt.cudaFreeHost(p)
