"""Measurement aid: host-to-device copy rate from ordinary pinned memory against write-combined pinned memory
(cudaHostAllocWriteCombined), 1 GiB per copy, through the CUDA runtime via ctypes."""
import ctypes as C
import torch

torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
n = 1 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
for name, flags in (("default", 0), ("write-combined", 4), ("portable+mapped", 3)):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        a.record(stream)
        for _ in range(5):
            assert rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(stream.cuda_stream)) == 0
        b.record(stream)
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%s: %.2f ms per GiB = %.1f GB/s" % (name, ms, n / ms / 1e6))
    rt.cudaFreeHost(p)
