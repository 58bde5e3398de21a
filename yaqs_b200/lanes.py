"""Concurrent lanes: several lock-step sub-batches of trajectories advance at the same time on one GPU, one host thread and
one CUDA stream each.

Why: a lock-step batch alternates between throughput-bound kernels (the DMMA contractions, the streaming Householder
kernels) and latency-bound ones (the implicit-QL rotation chain of the eigen-solver, tiny per-trajectory reductions, the
host reads that close a Lanczos / Jacobi loop).  With two or more independent lanes the latency-bound phases of one lane
run underneath the throughput-bound phases of another -- the GPU analogue of the reference keeping every core busy with
its own trajectory (core/parallel_utils.py:331-390).  ctypes calls and stream synchronisations release the GIL, so plain
Python threads are enough.
"""
from __future__ import annotations

import threading
from typing import Callable, Sequence

import torch


class Lanes:
    def __init__(self, device, n: int):
        self.device = torch.device(device)
        self.n = max(1, int(n))
        self.cuda = self.device.type == "cuda"
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.n)] if (self.cuda and self.n > 1) else []

    def map(self, fn: Callable, items: Sequence, serial: bool | str = False, between: Callable | None = None) -> list:
        """fn(item) for every item, at most ``n`` at a time, each on its own stream; results in item order.  ``serial``: one
        after the other on the caller's stream (a kernel timed with CUDA events must have the GPU to itself); ``serial="first"``:
        only the first item runs alone on the caller's stream, ``between()`` is called, the others run concurrently afterwards
        (bench.py's instrumented step: one lane under per-launch events, the rest of the step at full speed)."""
        items = list(items)
        if serial == "first" and self.streams and len(items) > 1:
            first = fn(items[0])
            if self.cuda:
                torch.cuda.current_stream(self.device).synchronize()
            if between is not None:
                between()
            return [first] + self.map(fn, items[1:], serial=False)
        if serial or not self.streams or len(items) <= 1:
            out = [fn(it) for it in items]
            if between is not None:
                between()
            return out
        out: list = [None] * len(items)
        errs: list = []
        main = torch.cuda.current_stream(self.device)
        for g0 in range(0, len(items), self.n):
            group = list(range(g0, min(len(items), g0 + self.n)))

            def work(slot: int, k: int) -> None:
                try:
                    torch.cuda.set_device(self.device)
                    st = self.streams[slot]
                    st.wait_stream(main)
                    with torch.cuda.stream(st):
                        out[k] = fn(items[k])
                        st.synchronize()
                except BaseException as e:          # noqa: BLE001 -- re-raised on the calling thread
                    errs.append(e)

            threads = [threading.Thread(target=work, args=(slot, k), daemon=True) for slot, k in enumerate(group)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errs:
                raise errs[0]
        return out
