"""Asynchronous sample pipeline in front of the forward: decode on worker threads, stage in pinned host memory, copy to the
device on a side stream while the previous sample is being computed (the reference feeds ``test_mvs4.py``
through a ``DataLoader`` with ``num_workers=4`` and a blocking ``tocuda``, ``test_mvs4.py:162-166,204``, ``utils.py:60-70``).

    for sample in Prefetcher(jobs, device="cuda:0"):          # jobs: callables returning one sample each
        out = model(sample["imgs"], sample["proj_matrices"], sample["depth_values"])

A sample is any nesting of dict / list / tuple whose leaves are numpy arrays or tensors (copied), or anything else (passed
through: file-name patterns, scalars).  Order is preserved.  At most ``depth`` samples are resident on the device ahead of
the consumer and ``depth + workers`` decoded on the host; an exception in a job is re-raised at the position of its sample.
On a CPU device the tensors are handed over as they are (host-side logic of the pipeline; used by the CPU tests).
"""
from __future__ import annotations

import collections
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Callable, Deque, Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


def map_leaves(x: Any, fn: Callable[[Tensor], Any]) -> Any:
    """Apply ``fn`` to every array / tensor leaf of a nested sample; other leaves are returned unchanged."""
    if isinstance(x, np.ndarray):
        return fn(torch.from_numpy(np.ascontiguousarray(x)))
    if isinstance(x, Tensor):
        return fn(x)
    if isinstance(x, dict):
        return {k: map_leaves(v, fn) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(map_leaves(v, fn) for v in x)
    return x


def batch_eval_sample(sample: Dict) -> Dict:
    """``formats.load_eval_sample`` output -> the batched (B = 1) tensors ``MVS4net.forward`` takes, still on the host."""
    return {"imgs": [torch.from_numpy(np.ascontiguousarray(i))[None] for i in sample["imgs"]],
            "proj_matrices": {k: torch.from_numpy(v)[None] for k, v in sample["proj_matrices"].items()},
            "depth_values": torch.from_numpy(sample["depth_values"])[None],
            "filename": [sample["filename"]]}


class _PinnedRing:
    """``depth`` generations of pinned staging buffers, reused when shapes repeat (they do: one scan = one frame size)."""

    def __init__(self, depth: int):
        self.slots: List[Dict[Tuple, List[Tensor]]] = [dict() for _ in range(depth)]
        self.events: List[Optional[torch.cuda.Event]] = [None] * depth
        self.cursor: List[Dict[Tuple, int]] = [dict() for _ in range(depth)]

    def begin(self, slot: int) -> None:
        if self.events[slot] is not None:
            self.events[slot].synchronize()   # the copy that last read this generation's buffers has finished
        self.cursor[slot] = {}

    def stage(self, slot: int, t: Tensor) -> Tensor:
        key = (tuple(t.shape), t.dtype)
        pool = self.slots[slot].setdefault(key, [])
        i = self.cursor[slot].get(key, 0)
        self.cursor[slot][key] = i + 1
        if i == len(pool):
            pool.append(torch.empty(t.shape, dtype=t.dtype, pin_memory=True))
        pool[i].copy_(t)
        return pool[i]


class Prefetcher:
    def __init__(self, jobs: Iterable[Callable[[], Any]], device="cuda:0", depth: int = 2, workers: int = 2,
                 transform: Optional[Callable[[Any], Any]] = None):
        if depth < 1 or workers < 1:
            raise ValueError("Prefetcher: depth and workers must be >= 1")
        self.jobs = jobs
        self.device = torch.device(device)
        self.depth, self.workers = depth, workers
        self.transform = transform
        self.h2d_bytes = 0          # bytes copied so far (what bench-style accounting needs)

    # -- one sample: host structure -> device structure (+ the event the consumer's stream has to wait for)
    def _upload(self, host: Any, slot: int, ring: Optional[_PinnedRing], stream) -> Tuple[Any, Optional[torch.cuda.Event]]:
        if self.device.type != "cuda":
            return map_leaves(host, lambda t: t), None
        ring.begin(slot)
        with torch.cuda.stream(stream):
            def up(t: Tensor) -> Tensor:
                self.h2d_bytes += t.numel() * t.element_size()
                return ring.stage(slot, t).to(self.device, non_blocking=True)
            dev = map_leaves(host, up)
            ev = torch.cuda.Event()
            ev.record(stream)
        ring.events[slot] = ev
        return dev, ev

    def __iter__(self) -> Iterator[Any]:
        cuda = self.device.type == "cuda"
        ring = _PinnedRing(self.depth) if cuda else None
        stream = torch.cuda.Stream(self.device) if cuda else None
        it = iter(self.jobs)
        decoding: Deque[Future] = collections.deque()
        resident: Deque[Tuple[Any, Optional[torch.cuda.Event]]] = collections.deque()
        n_up = 0

        def run(job):
            s = job()
            return self.transform(s) if self.transform else s

        with ThreadPoolExecutor(self.workers, thread_name_prefix="mvster-prefetch") as pool:
            def top_up() -> None:
                while len(decoding) < self.depth + self.workers:
                    try:
                        job = next(it)
                    except StopIteration:
                        return
                    decoding.append(pool.submit(run, job))

            try:
                top_up()
                while decoding or resident:
                    while decoding and len(resident) < self.depth:
                        host = decoding.popleft().result()      # re-raises a job's exception here, in order
                        top_up()
                        resident.append(self._upload(host, n_up % self.depth, ring, stream))
                        n_up += 1
                    dev, ev = resident.popleft()
                    if ev is not None:
                        cur = torch.cuda.current_stream(self.device)
                        cur.wait_event(ev)
                        map_leaves(dev, lambda t: t.record_stream(cur))   # allocated on the copy stream, consumed on this one
                    yield dev
            finally:
                for f in decoding:
                    f.cancel()


def eval_jobs(datapath: str, scan: str, pairs, nviews: int, **kw) -> List[Callable[[], Dict]]:
    """One job per (reference view, source views) entry of a pair file (``formats.read_pair_file``), producing the batched
    host tensors of that sample (``formats.load_eval_sample`` + ``batch_eval_sample``)."""
    from . import formats
    return [(lambda r=ref, s=tuple(srcs): batch_eval_sample(formats.load_eval_sample(datapath, scan, r, s, nviews, **kw)))
            for ref, srcs in pairs]
