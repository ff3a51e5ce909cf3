"""Host-to-device pipelining for batched matching.

``LightGlue.forward_async`` enqueues a forward without waiting for the GPU (its only host dependency, the
``stop`` / match-count words, is read back asynchronously), so the H2D copy of batch i+1 (pinned host
memory, separate copy stream) overlaps the kernels of batch i and the kernels of batch i+1 are queued
before batch i's result is handed out: the GPU never idles between batches.  Results come back in pinned
host tensors.  This is the throughput path a server fed from host memory uses (the
reference has no equivalent: its ``match_pair`` moves one pair at a time, utils.py:150-165).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator

import torch

RESULT_KEYS = ("matches0", "matches1", "matching_scores0", "matching_scores1")


def _shapes(batch: dict):
    return tuple((k, kk, tuple(vv.shape), vv.dtype) for k, v in sorted(batch.items()) for kk, vv in sorted(v.items()))


def match_stream(matcher, batches: Iterable[dict], device: torch.device | None = None) -> Iterator[Dict[str, torch.Tensor]]:
    """Yield one result dict (pinned CPU tensors: matches0/1, matching_scores0/1, plus ``stop``) per host batch.

    `batches` yields dicts in the matcher's input format whose tensors live in (ideally pinned) host
    memory.  Copies run on a side stream one batch ahead of the compute stream, into a ring of three persistent
    device input slots (no per-batch device allocation: a fresh 135 MB allocation per batch on the copy stream made
    the caching allocator fall back to cudaMalloc -- a device synchronisation -- whenever the previous batches' blocks
    were still held by the compute stream).  The yielded tensors live in a ring of three pinned buffers: a result
    stays valid until two further results have been yielded."""
    device = device or next(matcher.parameters()).device
    compute = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)
    it = iter(batches)
    ring = getattr(matcher, "_stream_inputs", None)
    if ring is None:
        ring = matcher._stream_inputs = {"sig": None, "slots": [], "free": []}
    n_staged = 0

    def stage(b):
        nonlocal n_staged
        sig = _shapes(b)
        if ring["sig"] != sig:  # (re)build the device input ring for this batch geometry
            ring["sig"] = sig
            ring["slots"] = [{k: {kk: torch.empty(vv.shape, dtype=vv.dtype, device=device) for kk, vv in v.items()} for k, v in b.items()}
                             for _ in range(3)]
            ring["free"] = [None, None, None]
        i = n_staged % 3
        n_staged += 1
        dev = ring["slots"][i]
        with torch.cuda.stream(copy):
            if ring["free"][i] is not None:
                copy.wait_event(ring["free"][i])  # the forward that last read this slot has finished with it
            for k, v in b.items():
                for kk, vv in v.items():
                    dev[k][kk].copy_(vv, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return dev, ev, i

    try:
        nxt = stage(next(it))
    except StopIteration:
        return
    pending = None  # (host result dict, event) of the previous batch
    # ring of pinned result buffers, kept on the matcher: pinned allocation (cudaHostAlloc) is slow, reuse it
    slots = getattr(matcher, "_stream_slots", None)
    if slots is None:
        slots = [dict(), dict(), dict()]
        matcher._stream_slots = slots
    n_done = 0
    while nxt is not None:
        dev, ev, slot_i = nxt
        try:
            nxt = stage(next(it))  # H2D of the following batch overlaps this batch's kernels
        except StopIteration:
            nxt = None
        compute.wait_event(ev)
        pend = matcher.forward_async(dev)  # queued behind the previous batch's kernels; no host wait
        freed = torch.cuda.Event()
        freed.record(compute)
        ring["free"][slot_i] = freed
        out = pend.tensors
        slot = slots[n_done % len(slots)]
        if not slot or any(slot[k].shape != out[k].shape for k in RESULT_KEYS):
            slot.clear()
            slot.update({k: torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True) for k in RESULT_KEYS})
        host = dict(slot)
        n_done += 1
        for k in RESULT_KEYS:
            host[k].copy_(out[k], non_blocking=True)
        done = torch.cuda.Event()
        done.record(compute)
        if pending is not None:
            yield _resolve(pending)
        pending = (host, done, pend)
    if pending is not None:
        yield _resolve(pending)


def _resolve(pending):
    host, done, pend = pending
    done.synchronize()            # the result copies of that batch have landed in the pinned slot
    host["stop"] = pend.result()["stop"]
    return host
