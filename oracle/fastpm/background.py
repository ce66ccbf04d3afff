"""Restatement of ``fastpm.background.MatterDominated`` (fastpm-python, un-pinned
third-party dependency of ``/root/reference/vmad/lib/fastpm.py:5,504,538``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Published algorithm (FastPM, Feng et al. 2016, and fastpm-python's
``background.py``): integrate the linear (D1) and second-order (D2) growth ODEs
of a flat LCDM background in ``ln a`` from deep matter domination, normalise
``D1(a=1) = 1`` and ``D2 / D1(1)^2``, then expose

    E, D1, D2, f1, f2            (used at fastpm.py:481-485, 566-570)
    Gp = D1,       gp = dD1/da   (drift factor, fastpm.py:443-444)
    Gf = a^3 E dD1/da, gf = dGf/da   (kick factor, fastpm.py:440-441)

With ' = d/dlna:   D1'' + (2 + dlnE/dlna) D1' = 3/2 Om(a) D1
                   D2'' + (2 + dlnE/dlna) D2' = 3/2 Om(a) (D2 - D1^2)
started at a0 = 1e-7 on the matter-dominated solution D1 = a, D2 = -3/7 a^2.
fastpm-python integrates with scipy ``odeint`` at default tolerance (~1e-8) and
interpolates linearly in ``ln a`` between the support points; here the
tolerance is tightened (rtol 1e-13, no absolute floor: D2 starts at 1e-14) so that the oracle and the product (which use
different integrators) agree far below the 1e-10 parity bar.
"""
import numpy as np
from scipy.integrate import odeint


class MatterDominated(object):
    def __vmad_hyper_key__(self):
        """memo key when the repo's VM memoises autooperator models"""
        return (self.Omega0_m, self.Omega0_lambda, self.Omega0_k, self.lna.tobytes())

    def __init__(self, Omega0_m, Omega0_lambda=None, Omega0_k=0, a=None, a_normalize=1.0):
        if Omega0_lambda is None:
            Omega0_lambda = 1 - Omega0_k - Omega0_m
        self.Omega0_m = Omega0_m
        self.Omega0_lambda = Omega0_lambda
        self.Omega0_k = Omega0_k
        self.Om0 = Omega0_m

        if a is None:
            lna = np.log(np.logspace(-7, 0, 1024 * 10, endpoint=True))
        else:
            a = np.array(a, dtype='f8', copy=True).ravel()
            a.sort()
            if a[0] > 1e-7:
                a = np.concatenate([[1e-7], a])
            if a[-1] < 1.0:
                a = np.concatenate([a, [1.0]])
            lna = np.log(a)
        self.lna = lna

        a0 = np.exp(lna[0])
        y0 = [a0, a0, -3. / 7 * a0 ** 2, -6. / 7 * a0 ** 2]
        y = odeint(self._ode, y0, lna, tfirst=False, rtol=1e-13, atol=1e-30, mxstep=100000)

        v1 = [y[:, 0], y[:, 1]]
        v1.append(np.array([self._ode(yi, li)[1] for yi, li in zip(y, lna)]))
        v2 = [y[:, 2], y[:, 3]]
        v2.append(np.array([self._ode(yi, li)[3] for yi, li in zip(y, lna)]))

        lnan = np.log(a_normalize)
        n1 = np.interp(lnan, lna, v1[0])
        self._D1 = [v / n1 for v in v1]
        self._D2 = [v / n1 ** 2 for v in v2]

    # -- background ----------------------------------------------------------
    def efunc(self, a):
        return (self.Omega0_m / a ** 3 + self.Omega0_k / a ** 2 + self.Omega0_lambda) ** 0.5

    def efunc_prime(self, a):
        """dE/da"""
        return 0.5 / self.efunc(a) * (-3 * self.Omega0_m / a ** 4 - 2 * self.Omega0_k / a ** 3)

    def E(self, a, order=0):
        if order == 0:
            return self.efunc(a)
        return self.efunc_prime(a)

    def Om(self, a):
        return (self.Omega0_m / a ** 3) / self.efunc(a) ** 2

    def _ode(self, y, lna):
        D1, F1, D2, F2 = y
        a = np.exp(lna)
        hfac = 2.0 + a * self.efunc_prime(a) / self.efunc(a)
        om = self.Om(a)
        return [F1, -hfac * F1 + 1.5 * om * D1,
                F2, -hfac * F2 + 1.5 * om * (D2 - D1 ** 2)]

    # -- growth --------------------------------------------------------------
    def D1(self, a, order=0):
        return np.interp(np.log(a), self.lna, self._D1[order])

    def D2(self, a, order=0):
        return np.interp(np.log(a), self.lna, self._D2[order])

    def f1(self, a):
        return self.D1(a, order=1) / self.D1(a, order=0)

    def f2(self, a):
        return self.D2(a, order=1) / self.D2(a, order=0)

    def Gp(self, a):
        return self.D1(a)

    def gp(self, a):
        return self.D1(a, order=1) / a

    def Gf(self, a):
        return self.D1(a, 1) * a ** 2 * self.E(a)

    def gf(self, a):
        return (self.D1(a, 2) * a ** 2 * self.E(a)
                + self.D1(a, 1) * (a ** 3 * self.E(a, order=1) + 2 * a ** 2 * self.E(a))) / a
