import numpy as np, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import udacitympc_b200 as mp
from udacitympc_b200 import synth
st, cf = synth.line_problems(97)
xs, ys = synth.roadmap_windows(97)
with mp.MPC(device=0) as m:
    fit = mp.polyfit_batch(xs, ys, 3, mpc=m)
    st3 = synth.roadmap_problems(97, fit)
    for mode in (0, 1):
        m.set_solver_mode(mode, 24, 0)
        r = m.solve_batch(st, cf, want_traj=True); assert (r['status']==0).all()
        r = m.solve_batch(st3, fit, want_traj=True); assert (r['status']==0).all()
    r = m.closed_loop(st[:5], cf[:5], 3)
    ks, ka = synth.kinematic_inputs(100, H=7)
    mp.rollout_batch(ks, ka, 0.3, 2.0, mpc=m)
    mp.polyfit_batch(np.sort(np.random.rand(33,12)), np.random.rand(33,12), 5, mpc=m)
    mp.roadmap_reference_batch(np.column_stack([synth.roadmap_centerline()[:50], np.zeros(50), np.full(50, 10.0)]), synth.roadmap_centerline(), mpc=m)
with mp.MPC(device=0, N=10) as m:
    m.set_solver_mode(0, 24, 0)
    r = m.solve_batch(st3[:40], fit[:40], want_traj=True)
print('sanitizer workload ok')
# round 2: pipelined solves (export into tail contexts, tails on their own streams), the asynchronous host-buffer call,
# the multi entry (two handles on one device), the strided cooperative kernel on a compacted batch
import ctypes, torch
dev = torch.device("cuda", 0)
B = 1536
xs, ys = synth.roadmap_windows(B)
with mp.MPC(device=0) as m:
    fit = mp.polyfit_batch(xs, ys, 3, mpc=m)
    st3 = synth.roadmap_problems(B, fit)
    m.set_solver_mode(0, 0, 256)     # per-pass path from 256 problems on
    plain = m.solve_batch(st3, fit)
    m.set_batch_split(1)
    m.set_pipeline(2, 128)
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    st_d = torch.from_numpy(np.ascontiguousarray(st3.T)).to(dev); cf_d = torch.from_numpy(np.ascontiguousarray(fit.T)).to(dev)
    outs = [dict(out8=torch.zeros((8, B), dtype=torch.float64, device=dev), status=torch.full((B,), -7, dtype=torch.int32, device=dev),
                 iters=torch.zeros(B, dtype=torch.int32, device=dev)) for _ in range(4)]
    for i, o in enumerate(outs):
        m.solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, o["out8"].data_ptr(), 0, 0, o["status"].data_ptr(), o["iters"].data_ptr(), streams[i % 2].cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        assert (o["status"].cpu().numpy() == plain["status"]).all() and np.abs(o["out8"].T.cpu().numpy() - plain["out8"]).max() < 1e-9
    m.set_pipeline(0, 0)
lib = mp.load_library()
dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
hs = [mp.MPC(device=0) for _ in range(2)]
o8 = np.zeros((B, 8)); stt = np.zeros(B, dtype=np.int32)
harr = (ctypes.c_void_p * 2)(*[h.handle for h in hs])
assert lib.b200mpc_solve_batch_multi(harr, 2, B, st3.ctypes.data_as(dp), fit.ctypes.data_as(dp), 4, o8.ctypes.data_as(dp), None, None, stt.ctypes.data_as(ip), None) == 0
assert (stt == 0).all() and np.abs(o8 - plain["out8"]).max() < 1e-9
o8b = np.zeros((B, 8))
assert lib.b200mpc_solve_batch_async(hs[0].handle, B, st3.ctypes.data_as(dp), fit.ctypes.data_as(dp), 4, o8b.ctypes.data_as(dp), None, None, None, None) == 0
hs[0].wait()
assert np.abs(o8b - plain["out8"]).max() < 1e-9
for h in hs:
    h.close()
print('round-2 sanitizer workload ok')
