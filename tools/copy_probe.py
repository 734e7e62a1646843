#!/usr/bin/env python
"""Development aid: what the host link does with the C3 step's copies cut into pieces (pinned memory, CUDA events / host clock)."""
import time
import torch

dev = torch.device("cuda:0")
N = 10_000_000
h_in = [torch.empty(N, dtype=torch.int64).pin_memory(), torch.empty(N, dtype=torch.int64).pin_memory(), torch.empty(N, dtype=torch.int32).pin_memory()]
h_out = [torch.empty(N, dtype=torch.float64).pin_memory() for _ in range(3)]
d_in = [torch.empty_like(t, device=dev) for t in h_in]
d_out = [torch.empty_like(t, device=dev) for t in h_out]
s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
slots = [torch.cuda.Stream() for _ in range(3)]


def run(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print("%-60s %.3f ms" % (name, (time.perf_counter() - t0) / reps * 1e3))


def both(piece_in, piece_out, streams_out=None):
    def f():
        with torch.cuda.stream(s_h2d):
            for a in range(0, N, piece_in):
                for h, d in zip(h_in, d_in):
                    d[a:a + piece_in].copy_(h[a:a + piece_in], non_blocking=True)
        k = 0
        for a in range(0, N, piece_out):
            st = s_d2h if streams_out is None else streams_out[k % len(streams_out)]
            k += 1
            with torch.cuda.stream(st):
                for h, d in zip(h_out, d_out):
                    h[a:a + piece_out].copy_(d[a:a + piece_out], non_blocking=True)
    return f


def d2h_only(piece):
    def f():
        with torch.cuda.stream(s_d2h):
            for a in range(0, N, piece):
                for h, d in zip(h_out, d_out):
                    h[a:a + piece].copy_(d[a:a + piece], non_blocking=True)
    return f


def slot_pipeline(piece):
    """what mg_h_pipeline enqueues, without the kernels: per chunk 3 uploads + 3 downloads on the chunk's slot stream"""
    def f():
        k = 0
        for a in range(0, N, piece):
            st = slots[k % 3]
            k += 1
            with torch.cuda.stream(st):
                for h, d in zip(h_in, d_in):
                    d[a:a + piece].copy_(h[a:a + piece], non_blocking=True)
                for h, d in zip(h_out, d_out):
                    h[a:a + piece].copy_(d[a:a + piece], non_blocking=True)
    return f


run("D2H only, one piece", d2h_only(N))
run("D2H only, 393216-row pieces", d2h_only(393216))
run("D2H only, 131072-row pieces", d2h_only(131072))
run("both directions, one piece each", both(N, N))
run("both, 393216-row pieces, one stream per direction", both(393216, 393216))
run("both, 393216 in / 1572864 out", both(393216, 1572864))
run("both, 393216-row pieces, downloads over 3 streams", both(393216, 393216, slots))
for p in (262144, 393216, 524288, 1048576):
    run("slot pipeline (3 streams, up + down per chunk), %d rows" % p, slot_pipeline(p))
