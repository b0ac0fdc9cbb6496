"""Latency of one host-buffer solve call vs batch size, for the cooperative-only path and the per-pass path."""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import udacitympc_b200 as mp
from udacitympc_b200 import synth
st, cf = synth.line_problems(32768)
for name, args in (('coop-only', (0, 16, 1 << 30)), ('per-pass+coop', (0, 16, 0)), ('fused-thread', (1, 0, -1))):
    with mp.MPC(device=0) as m:
        m.set_solver_mode(*args)
        for B in (1, 256, 2048, 4096, 8192, 16384, 32768):
            if name == 'fused-thread' and B > 2048:
                continue
            m.solve_batch(st[:B], cf[:B])
            t = time.perf_counter(); n = 8
            for _ in range(n):
                m.solve_batch(st[:B], cf[:B])
            dt = (time.perf_counter() - t) / n
            print(f'{name:14s} B {B:6d} ms/call {dt*1e3:8.3f} solves/s {B/dt:12.0f}')
