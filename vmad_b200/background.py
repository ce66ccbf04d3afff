"""Growth factors of a flat LCDM background: the scalars the FastPM leapfrog needs.

Stands in for ``fastpm.background.MatterDominated`` (third-party, used at
vmad/lib/fastpm.py:5,504,538 and queried at :441,444,481-485): ``E, D1, D2, f1, f2, Gp, gp,
Gf, gf, Om0``.  Host-side scalar code, not a kernel: the growth ODEs

    D1'' + (2 + dlnE/dlna) D1' = 3/2 Om(a) D1
    D2'' + (2 + dlnE/dlna) D2' = 3/2 Om(a) (D2 - D1^2)          (' = d/dlna)

are integrated once from a = 1e-7 (matter-dominated start D1 = a, D2 = -3/7 a^2) with an
explicit 8th-order Runge-Kutta (scipy DOP853) to 1e-13 and evaluated exactly at the scale
factors requested at construction (the leapfrog's stage and mid-stage times); other arguments
are interpolated linearly in ln a like fastpm-python does.  D1 is normalised to 1 and D2 to
D1(1)^2 at a = 1.
"""
import numpy as np
from scipy.integrate import solve_ivp


class MatterDominated(object):
    def __init__(self, Omega0_m, Omega0_lambda=None, Omega0_k=0, a=None, a_normalize=1.0):
        if Omega0_lambda is None:
            Omega0_lambda = 1 - Omega0_k - Omega0_m
        self.Omega0_m = Omega0_m
        self.Omega0_lambda = Omega0_lambda
        self.Omega0_k = Omega0_k
        self.Om0 = Omega0_m

        if a is None:
            grid = np.logspace(-7, 0, 10241)
        else:
            grid = np.unique(np.concatenate([[1e-7], np.asarray(a, dtype='f8').ravel(),
                                             [a_normalize, 1.0]]))
        self.lna = np.log(grid)
        a0 = grid[0]
        y0 = [a0, a0, -3. / 7 * a0 ** 2, -6. / 7 * a0 ** 2]
        sol = solve_ivp(self._rhs, (self.lna[0], self.lna[-1]), y0, method='DOP853',
                        t_eval=self.lna, rtol=1e-13, atol=1e-30)
        if not sol.success:
            raise RuntimeError("growth ODE integration failed: %s" % sol.message)
        y = sol.y
        acc = np.array([self._rhs(t, yi) for t, yi in zip(self.lna, y.T)])
        n1 = np.interp(np.log(a_normalize), self.lna, y[0])
        self._D1 = [y[0] / n1, y[1] / n1, acc[:, 1] / n1]
        self._D2 = [y[2] / n1 ** 2, y[3] / n1 ** 2, acc[:, 3] / n1 ** 2]

    def __vmad_hyper_key__(self):
        """memo key as a hyper-argument of an autooperator (vm/ops.py:_hyper_key)"""
        return (self.Omega0_m, self.Omega0_lambda, self.Omega0_k, self.lna.tobytes())

    # -- background expansion ---------------------------------------------------------------
    def efunc(self, a):
        return (self.Omega0_m / a ** 3 + self.Omega0_k / a ** 2 + self.Omega0_lambda) ** 0.5

    def efunc_prime(self, a):
        return 0.5 / self.efunc(a) * (-3 * self.Omega0_m / a ** 4 - 2 * self.Omega0_k / a ** 3)

    def E(self, a, order=0):
        return self.efunc(a) if order == 0 else self.efunc_prime(a)

    def Om(self, a):
        return (self.Omega0_m / a ** 3) / self.efunc(a) ** 2

    def _rhs(self, lna, y):
        D1, F1, D2, F2 = y
        a = np.exp(lna)
        drag = 2.0 + a * self.efunc_prime(a) / self.efunc(a)
        om = self.Om(a)
        return [F1, 1.5 * om * D1 - drag * F1, F2, 1.5 * om * (D2 - D1 * D1) - drag * F2]

    # -- growth --------------------------------------------------------------------------------
    def D1(self, a, order=0):
        return np.interp(np.log(a), self.lna, self._D1[order])

    def D2(self, a, order=0):
        return np.interp(np.log(a), self.lna, self._D2[order])

    def f1(self, a):
        return self.D1(a, 1) / self.D1(a, 0)

    def f2(self, a):
        return self.D2(a, 1) / self.D2(a, 0)

    def Gp(self, a):
        return self.D1(a)

    def gp(self, a):
        return self.D1(a, 1) / a

    def Gf(self, a):
        return self.D1(a, 1) * a ** 2 * self.E(a)

    def gf(self, a):
        E, dE = self.E(a), self.E(a, order=1)
        return (self.D1(a, 2) * a ** 2 * E + self.D1(a, 1) * (a ** 3 * dE + 2 * a ** 2 * E)) / a
