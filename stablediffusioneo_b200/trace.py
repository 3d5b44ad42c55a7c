"""Kernel timeline of the library's own launches (sdeo_set_trace): block (0,0,0) of every libsdeo kernel records the
start of the block, the moment its grid dependency resolved (programmatic dependent launch) and the end of the block on
the GPU's globaltimer. Profiling aid for bench.py / tools/step_timeline.py; costs one branch per kernel when off."""
import collections
import ctypes

import torch

from . import _lib

KINDS = {1: "conv", 2: "attention", 3: "groupnorm", 4: "layernorm", 5: "elementwise"}
Record = collections.namedtuple("Record", "kind grid mode splits bn start dep end halo", defaults=(0,))


def capture(fn, device, capacity=8192):
    """Runs fn() with tracing on and returns the records (times in microseconds from the first start)."""
    buf = torch.zeros(4 + 4 * capacity, dtype=torch.int64, device=device)
    buf[1] = capacity
    lib = _lib.load()
    torch.cuda.synchronize(device)
    _lib.check(lib.sdeo_set_trace(ctypes.c_void_p(buf.data_ptr())), "set_trace")
    try:
        fn()
        torch.cuda.synchronize(device)
    finally:
        lib.sdeo_set_trace(None)
    b = buf.cpu()
    n = min(int(b[0]), capacity)
    rec = b[4:4 + 4 * n].reshape(n, 4)
    if n == 0:
        return []
    t0 = int(rec[:, 1].min())
    out = []
    for i in range(n):
        tag = int(rec[i, 0])
        out.append(Record(KINDS.get(tag & 0xFF, "?"), (tag >> 8) & 0xFFFFFFFF, (tag >> 40) & 0xF, (tag >> 44) & 0xF,
                          (tag >> 48) & 0xFFF, (int(rec[i, 1]) - t0) / 1e3, (int(rec[i, 2]) - t0) / 1e3,
                          (int(rec[i, 3]) - t0) / 1e3, (tag >> 60) & 1))
    return out


def summarize(records):
    """Per kind: launches, busy = sum(dependency resolved -> block-0 end), tail = idle time that FOLLOWS a kernel of
    this kind until the next kernel (of any stream) is past its dependency (the rest of the grid finishing, the write
    drain and the dependent-launch latency: time the step pays for that launch without any kernel making progress).
    Also the busy union and the span of the whole trace."""
    per = collections.OrderedDict()
    for r in records:
        a = per.setdefault(r.kind, {"launches": 0, "busy_us": 0.0, "tail_us": 0.0, "hidden_prologue_us": 0.0})
        a["launches"] += 1
        a["busy_us"] += r.end - r.dep
        a["hidden_prologue_us"] += r.dep - r.start
    cur_end, last, union = None, None, 0.0
    for r in sorted(records, key=lambda r: r.dep):
        if cur_end is not None and r.dep > cur_end:
            per[last.kind]["tail_us"] += r.dep - cur_end
        if cur_end is None or r.end > cur_end:
            union += r.end - (max(cur_end, r.dep) if cur_end is not None else r.dep)
            cur_end, last = r.end, r
    span = max(r.end for r in records) if records else 0.0
    # per kind: wall-clock time during which at least one kernel of that kind is past its dependency (launches of the
    # two streams that run side by side are not counted twice)
    for kind, a in per.items():
        cur_e, u = None, 0.0
        for r in sorted((r for r in records if r.kind == kind), key=lambda r: r.dep):
            if cur_e is None or r.dep > cur_e:
                u += r.end - r.dep
                cur_e = r.end
            elif r.end > cur_e:
                u += r.end - cur_e
                cur_e = r.end
        a["union_us"] = u
    return {"kinds": per, "busy_union_us": union, "span_us": span}
