"""Development aid: run the LUW_TRACE build (make -C latticeurbanwind_b200/csrc OUT=../lib_trace EXTRA=-DLUW_TRACE=10) and print the per-tile
timeline of one CTA. usage: LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py [precision] [feat]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latticeurbanwind_b200 import cases, _cabi as A
from latticeurbanwind_b200.domain import Domain
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
feat = int(sys.argv[2]) if len(sys.argv) > 2 else 0
Nx, Ny, Nz = 512, 512, 256
case = sys.argv[3] if len(sys.argv) > 3 else "periodic_box"
zones = dict(downstream_face=2, buffer_N=16, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=20, sponge_inv_tau=0.02)
with Domain(Nx, Ny, Nz, precision=prec, features=feat, w=cases.relaxation_rate(1e-6 if feat & 8 else 1 / 6), arith=1, **zones) as d:
    cases.block_case(case, (Nx, Ny, Nz), out=(d.flags, d.rho, d.u))
    d.omega = (0, 5.6e-6, 4.7e-6)
    d.upload_all(); d.t = 1; d.enqueue_initialize(); d.t = 0
    d.run_steps(5); d.finish_queue()
    d.timer_begin(); d.run_steps(1); ms = d.timer_end()
    buf = np.zeros((6, 2048), np.int64)
    rc = A.lib().luw_debug_trace(buf.ctypes.data_as(C.c_void_p))
    print("rc", rc, "step ms", ms)
    ns, cyc = buf[2][2047] - buf[0][2047], buf[3][2047] - buf[1][2047]
    print(f"traced CTA alive {ns/1e6:.4f} ms = {cyc} cycles -> SM clock {cyc/max(ns,1)*1e3:.0f} MHz")
    t0 = buf[4][0]
    n = int((buf[4] > 0).sum())
    print("tiles traced:", n)
    names = ["p_done", "p_commit", "p_read", "p_loads", "c_start", "c_end"]
    rel = (buf[:, :n] - t0)
    if os.environ.get("TRACE_ROWS"):
        for q in list(range(0, 24)) + list(range(n - 6, n)):
            print(q, " ".join(f"{names[k]}={rel[k][q]:8d}" for k in (4, 5, 0, 1, 2, 3)))
    per_tile = np.diff(buf[5][:n])
    T = int(os.environ.get("TRACE_TILES_X", "0"))
    if T:
        pt = per_tile[T * 4 - 1:]  # skip the start-up strips; pt[j] = period ending at tile j+T*4
        m = (len(pt) // T) * T
        print("mean period by position in strip:", np.round(pt[:m].reshape(-1, T).mean(axis=0)).astype(int).tolist())
    print("period percentiles 10/50/90/99:", np.percentile(per_tile[32:], [10, 50, 90, 99]).astype(int).tolist(), "steady mean", per_tile[32:].mean())
    print("consumer tile period: median", np.median(per_tile), "mean", per_tile.mean())
    print("compute (c_end-c_start): median", np.median(buf[5][:n] - buf[4][:n]))
    print("producer: done->commit", np.median(buf[1][:n] - buf[0][:n]), "commit->read", np.median((buf[2] - buf[1])[:n - 8]), "read->loads", np.median((buf[3] - buf[2])[:n - 8]))
    print("consumer end -> producer sees done", np.median(buf[0][:n] - buf[5][:n]))
