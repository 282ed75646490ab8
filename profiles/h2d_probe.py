#!/usr/bin/env python
"""Host-to-device copy ceiling of this platform: pinned (default) against write-combined pinned memory, copy sizes from
32 MB to 1 GB, one and two streams.  CUDA-event times, GB/s.  ("try cudaHostAllocWriteCombined and >= 128 MB
copies before declaring the ceiling".)   python profiles/h2d_probe.py"""
import ctypes
import json

import torch

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
H2D = 1


def alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags)
    assert rc == 0, rc
    ctypes.memset(p, 1, nbytes)         # touch every page
    return p


def main():
    torch.cuda.init()
    total = 1 << 30
    dev = torch.empty(total, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(2)]
    for kind, flags in (("pinned", 0), ("pinned write-combined", 4)):
        host = alloc(total, flags)
        for chunk_mb in (32, 64, 128, 256, 1024):
            for ns in (1, 2):
                chunk = chunk_mb << 20
                n = total // chunk
                if ns > n:
                    continue
                best = 0.0
                for _ in range(3):
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for s in streams[:ns]:
                        s.wait_event(e0)
                    for i in range(n):
                        s = streams[i % ns]
                        rc = rt.cudaMemcpyAsync(dev.data_ptr() + i * chunk, host.value + i * chunk, chunk, H2D, s.cuda_stream)
                        assert rc == 0, rc
                    for s in streams[:ns]:
                        torch.cuda.current_stream().wait_stream(s)
                    e1.record()
                    torch.cuda.synchronize()
                    best = max(best, total / e0.elapsed_time(e1) / 1e6)
                print(json.dumps({"host_memory": kind, "chunk_MB": chunk_mb, "streams": ns, "h2d_GBps": round(best, 2)}), flush=True)
        rt.cudaFreeHost(host)


if __name__ == "__main__":
    main()
