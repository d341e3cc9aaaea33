"""Does overlapping the sweeps of two half batches hide the last-round tail of the generated-problem kernels?
cfg3 (100k candidates of the 6-state problem): one handle with the whole batch against two handles with half of it each
on two CUDA streams (the block scheduler fills the SMs a kernel's last round leaves idle with the other half's blocks)."""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
ocs = importlib.import_module("optimal-control-solvers_b200")
dev = torch.device("cuda", 0)


def setup(Bt, stream):
    wl = bench.make_workload("cfg3", ocs, 0, Bt)
    integ = wl["integ"]
    integ.attach_control(wl["ctl"])
    integ.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev); g.manual_seed(2)
    x0 = (1.0 + 1.5 * torch.rand((6, Bt), dtype=torch.float64, device=dev, generator=g)).contiguous()
    v = 0.05 * torch.randn(wl["v_batch"], dtype=torch.float64, device=dev, generator=g)
    v[:2] = 0.2 + 0.6 * torch.rand((2, Bt), dtype=torch.float64, device=dev, generator=g)
    out = dict(J=torch.empty(Bt, dtype=torch.float64, device=dev), xT=torch.empty((6, Bt), dtype=torch.float64, device=dev),
               st=torch.empty(Bt, dtype=torch.int32, device=dev), g=torch.empty((wl["rows"], Bt), dtype=torch.float64, device=dev),
               l0=torch.empty((7, Bt), dtype=torch.float64, device=dev))
    return wl, integ, x0, v.contiguous(), out


def ev(wl, integ, x0, v, out, first):
    integ.evaluate_dev(wl["prob"], x0.data_ptr(), x0.shape[1], v=None, v_ptr=v.data_ptr(), J_ptr=out["J"].data_ptr(),
                       xT_ptr=out["xT"].data_ptr(), status_ptr=out["st"].data_ptr(), dJdv_ptr=out["g"].data_ptr(),
                       lam0_ptr=out["l0"].data_ptr(), bind=first)


def timed(sets, steps=10):
    for k in range(3):
        for s in sets:
            ev(*s, first=(k == 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream(dev))
    for st in streams:
        st.wait_event(e0)
    for _ in range(steps):
        for s in sets:
            ev(*s, first=False)
    for st in streams:
        torch.cuda.default_stream(dev).wait_stream(st)
    e1.record(torch.cuda.default_stream(dev))
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
Bt = 100_000
res = {}
one = setup(Bt, streams[0])
res["one_handle_ms"] = timed([one])
for split in ((50_048, 49_952), (75_776, 24_224), (66_688, 33_312)):
    a, b = setup(split[0], streams[0]), setup(split[1], streams[1])
    res["two_handles_%d_%d_ms" % split] = timed([a, b])
    a[1].close(); b[1].close()
print(json.dumps(res))
