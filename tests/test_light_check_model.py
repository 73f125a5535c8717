"""CPU model of the bookkeeping behind the light check: while the checks are dense the passes carry
axsum = A xSum and atysum = A'ySum next to xSum / ySum, with the same deferred weights, so that a check can form
A xbar = axsum / sum(w) and A'ybar = atysum / sum(w) without an SpMV.  The model replays the update rules of the kernels
(K1 / K2 / K3 flushes with the pending weight, rejected passes, the check's own flush, restarts to the average and to the
current iterate) on random data and checks the invariant after every event -- the CUDA kernels are compared with the SpMV
variant on hardware (tests/test_gpu_solve.py::test_light_check_matches_the_spmv_check); this file pins the RULES."""
import numpy as np
import pytest


class Model:
    def __init__(self, A, x0, y0, xsum0):
        self.A = A
        self.x, self.y = x0.copy(), y0.copy()
        self.ax, self.aty = A @ self.x, A.T @ self.y
        self.xsum, self.ysum = xsum0.copy(), np.zeros_like(y0)           # sums start at proj(0) / 0 (PDHG_Init_Variables)
        self.axsum, self.atysum = A @ self.xsum, np.zeros_like(x0)      # engine.cu: copy of ax[0] (cold) or one SpMV (hot start)
        self.pending, self.w, self.sum_step = False, 0.0, 0.0

    def flush(self):
        """what K1/K2/K3 (or a check's sweeps) do with the pending weight of the accepted iterate"""
        if self.pending:
            self.xsum += self.w * self.x
            self.ysum += self.w * self.y
            self.axsum += self.w * self.ax          # K2: ax of the accepted iterate is in registers
            self.atysum += self.w * self.aty        # K3: aty of the accepted iterate is in registers
            self.pending = False

    def step(self, rng, accept):
        self.flush()                                 # the next pass adds the previous pass's weight first
        xn, yn = rng.standard_normal(self.x.size), rng.standard_normal(self.y.size)
        if accept:                                   # K4: the trial becomes current, its weight is pending
            self.x, self.y = xn, yn
            self.ax, self.aty = self.A @ xn, self.A.T @ yn
            self.w = float(rng.random()) + 0.1
            self.sum_step += self.w
            self.pending = True

    def check(self):
        self.flush()                                 # C1 / the sweeps flush the pending weight exactly once
        scale = 1.0 / self.sum_step if self.sum_step > 0 else 1.0
        return self.xsum * scale, self.ysum * scale, self.axsum * scale, self.atysum * scale

    def restart(self, to_average):
        xa, ya, axa, atya = self.check()
        if to_average:
            self.x, self.y, self.ax, self.aty = xa, ya, axa, atya     # restart_sweep: current <- average, products included
        self.xsum[:] = 0; self.ysum[:] = 0; self.axsum[:] = 0; self.atysum[:] = 0
        self.sum_step = 0.0

    def invariant(self):
        tol = 1e-10 * (1 + np.abs(self.axsum).max() + np.abs(self.atysum).max())
        return (np.abs(self.A @ self.xsum - self.axsum).max() <= tol and np.abs(self.A.T @ self.ysum - self.atysum).max() <= tol
                and np.abs(self.A @ self.x - self.ax).max() <= 1e-10 * (1 + np.abs(self.ax).max()))


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("hot", [False, True])
def test_carried_products_follow_the_sums(seed, hot):
    rng = np.random.default_rng(seed)
    m, n = 40, 55
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.2)
    proj0 = np.where(rng.random(n) < 0.3, 0.25, 0.0)            # columns with lower > 0: proj(0) != 0
    x0 = rng.standard_normal(n) if hot else proj0.copy()       # hot start: x0 != proj(0)
    M = Model(A, x0, np.zeros(m) if not hot else rng.standard_normal(m), proj0)
    assert M.invariant()
    for it in range(60):
        ev = rng.random()
        if ev < 0.55:
            M.step(rng, accept=True)
        elif ev < 0.7:
            M.step(rng, accept=False)               # rejected line-search step: nothing becomes pending
        elif ev < 0.9:
            xa, ya, axa, atya = M.check()
            assert np.allclose(A @ xa, axa, rtol=1e-9, atol=1e-9) and np.allclose(A.T @ ya, atya, rtol=1e-9, atol=1e-9)
        else:
            M.restart(to_average=bool(rng.random() < 0.5))
        assert M.invariant(), it
