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
