"""Development aid: where does the end-to-end loop of bench.py lose time? Variants of the per-step host loop on channel512_fp16s."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from latticeurbanwind_b200 import _cabi as A, cases
from latticeurbanwind_b200.domain import Domain, CellSet, pinned_empty
d = bench.build_domain(Domain, cases, "channel512_fp16s", A.ARITH_FAST, 0, pinned=pinned_empty)
d.upload_all(); d.t = 1; d.enqueue_initialize(); d.t = 0; d.finish_queue()
d.run_steps(20); d.finish_queue()
Nx, Ny, Nz = 512, 512, 512
yz = (np.arange(Ny, dtype=np.uint64)[None, :] + np.arange(Nz, dtype=np.uint64)[:, None] * np.uint64(Ny)).reshape(-1) * np.uint64(Nx)
inlet = yz[d.flags[yz.astype(np.int64)] == 2]
probe = yz + np.uint64(Nx - 2)
cin, cpr = CellSet(d, inlet), CellSet(d, probe)
uin = pinned_empty(3 * cin.count, np.float32); uin[:] = 0.05
upr, rpr = pinned_empty(3 * cpr.count, np.float32), pinned_empty(cpr.count, np.float32)
K = 100
def run(name, up, down, sync_every):
    d.finish_queue()
    t0 = time.perf_counter()
    for k in range(K):
        if up: cin.upload(A.FIELD_U, uin)
        d.enqueue_stream_collide(); d.increment_time_step()
        if down: cpr.download(A.FIELD_U, upr); cpr.download(A.FIELD_RHO, rpr)
        if sync_every and k % sync_every == sync_every - 1: d.finish_queue()
    d.finish_queue()
    dt = (time.perf_counter() - t0) / K * 1e3
    print(f"{name:40s} {dt:.3f} ms/step", flush=True)
d.timer_begin(); d.run_steps(K); print("batched run_steps (events)               %.3f ms/step" % (d.timer_end() / K))
run("steps only, sync at end", False, False, 0)
run("steps only, sync every 4", False, False, 4)
run("steps + upload, sync every 4", True, False, 4)
run("steps + downloads, sync every 4", False, True, 4)
run("steps + upload + downloads, sync every 4", True, True, 4)
run("steps + upload + downloads, sync at end", True, True, 0)
run("steps + upload + downloads, sync every step", True, True, 1)
d.timer_begin(); d.run_steps(K); print("batched run_steps (events)               %.3f ms/step" % (d.timer_end() / K))
