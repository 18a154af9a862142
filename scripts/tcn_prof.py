"""TCN forward timing probe (diagnostic): whole embed_clouds vs the C call alone vs per-kernel warm times (kineto)."""
import ctypes as C, json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from midastouch_b200._lib import call, ptr, stream_ptr
from midastouch_b200.tcn import pack_coordinates

dev = torch.device("cuda:0")
quick = bool(os.environ.get("TCN_NO_KINETO"))  # under ncu: a few forwards only
out = {"embed_clouds_ms": bench.tcn_time(dev, reps=2 if quick else 50)}
tcn, cloud = bench.tcn_time.last  # (set by tcn_time)


def ev_time(fn, reps=50):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return {"device_ms": e0.elapsed_time(e1) / reps, "host_issue_ms": (t1 - t0) * 1e3 / reps}


if quick:
    print(json.dumps(out))
    sys.exit(0)
out["embed_clouds"] = ev_time(lambda: tcn.embed_clouds(cloud))
if True:
    B, Pn, _ = cloud.shape
    ijk = torch.floor(cloud.reshape(-1, 3).float() / tcn.quantization_size).to(torch.int64)
    keys = torch.unique(pack_coordinates(torch.arange(B, device=dev).repeat_interleave(Pn), ijk))
    o = torch.empty((B, 256), dtype=torch.float64, device=dev)
    out["c_call_keys"] = ev_time(lambda: call("mt_tcn_forward", tcn._h, ptr(keys), keys.numel(), B, 1, ptr(o), 0, stream_ptr()))
try:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            tcn.embed_clouds(cloud)
        torch.cuda.synchronize()
    rows = [(e.key[:60], e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages()]
    rows.sort(key=lambda r: -r[1] * r[2])
    out["kernels_us"] = [(k, c // 10, round(t, 1)) for k, c, t in rows[:40]]
    out["kernels_total_us_per_forward"] = sum(c * t for _, c, t in rows) / 10
    evs = [e for e in prof.events() if e.device_time_total > 0 and "Memcpy" not in e.name and "Memset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 10
    last = evs[-per:]
    t0 = last[0].time_range.start
    out["sequence_last_forward"] = [(e.name[:28], round(e.time_range.start - t0, 1), round(e.device_time_total, 1)) for e in last]
except Exception as e:  # kineto may be unavailable on the box
    out["profiler_error"] = repr(e)
print(json.dumps(out, indent=1))
