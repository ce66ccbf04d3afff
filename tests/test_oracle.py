"""The oracle checked against what the reference's own tests pin for this path:
closed-form identities, finite-difference adjoint consistency, and its C loops against its
numpy loops.  (The reference's full test_fastpm.py runs on the oracle in
tests/test_reference_compat.py.)"""
import numpy
import pytest


def test_identities(oracle_pm):
    pm = oracle_pm([8, 8, 8], 16.0)
    x = pm.generate_whitenoise(seed=300, unitary=True, type='real')
    # test_fastpm.py:25-37: c2r(r2c(x)) == x
    assert numpy.allclose(x.r2c().c2r().value, x.value, rtol=0, atol=1e-13)
    # test_fastpm.py:158-159,176-177: a uniform one-per-cell grid paints exactly 1 per cell
    for shift in (0.0, 0.1, 0.5):
        q = pm.generate_uniform_particle_grid(shift=shift)
        assert numpy.abs(pm.paint(q, 1.0).value - 1.0).max() < 1e-13
    # test_fastpm.py:216-217: mass conservation
    rng = numpy.random.default_rng(1)
    pos = rng.uniform(-3, 20, size=(1000, 3))
    m = rng.uniform(0.5, 2, size=1000)
    assert abs(pm.paint(pos, m).value.sum() - m.sum()) < 1e-10
    # test_fastpm.py:142: Parseval, cdot(r2c x, r2c x1) = sum x x1 / N
    x1 = pm.generate_whitenoise(seed=301, unitary=True, type='real')
    lhs = x.r2c().cdot(x1.r2c()).real
    assert abs(lhs - (x.value * x1.value).sum() / pm.Ntot) < 1e-12


def test_layout_is_stable_sort_and_adjoint(oracle_pm):
    pm = oracle_pm([4, 6, 5], [4.0, 3.0, 10.0])
    rng = numpy.random.default_rng(2)
    pos = rng.uniform(-5, 15, size=(500, 3))
    lay = pm.decompose(pos)
    keys = pm.cell_keys(pos)
    assert numpy.array_equal(lay.indices, numpy.argsort(keys, kind='stable'))
    assert (numpy.diff(keys[lay.indices]) >= 0).all()
    a = rng.standard_normal((500, 3))
    b = rng.standard_normal((500, 3))
    # exchange / gather are mutual adjoints (fastpm.py:281-308)
    assert abs((lay.exchange(a) * b).sum() - (a * lay.gather(b)).sum()) < 1e-10
    assert numpy.array_equal(lay.gather(lay.exchange(a)), a)


def test_paint_readout_adjoint_and_gradients(oracle_pm):
    pm = oracle_pm([6, 6, 6], 12.0)
    rng = numpy.random.default_rng(3)
    pos = rng.uniform(0, 12, size=(300, 3))
    m = rng.uniform(0.5, 2, size=300)
    g = pm.create('real', value=rng.standard_normal(pm.rshape))
    # <paint(x, m), g> == <m, readout(g, x)>
    assert abs((pm.paint(pos, m).value * g.value).sum() - (m * g.readout(pos)).sum()) < 1e-10
    # paint_vjp / readout_vjp position gradients against central differences
    _pos, _mass = pm.paint_vjp(g, pos, mass=m)
    eps = 1e-6
    for (p, d) in [(0, 0), (7, 1), (123, 2)]:
        dp = numpy.zeros_like(pos)
        dp[p, d] = eps
        fd = ((pm.paint(pos + dp, m).value - pm.paint(pos - dp, m).value) * g.value).sum() / (2 * eps)
        assert abs(fd - _pos[p, d]) < 1e-6 * max(1.0, abs(fd))
    v = rng.standard_normal(300)
    _mesh, _x = g.readout_vjp(pos, v)
    for (p, d) in [(1, 0), (50, 1), (299, 2)]:
        dp = numpy.zeros_like(pos)
        dp[p, d] = eps
        fd = ((g.readout(pos + dp) - g.readout(pos - dp)) * v).sum() / (2 * eps)
        assert abs(fd - _x[p, d]) < 1e-6 * max(1.0, abs(fd))
    # jvp == directional derivative
    xdot = rng.standard_normal(pos.shape)
    j = pm.paint_jvp(pos, mass=m, v_pos=xdot).value
    fd = (pm.paint(pos + eps * xdot, m).value - pm.paint(pos - eps * xdot, m).value) / (2 * eps)
    assert numpy.abs(j - fd).max() < 1e-5


def test_c_loops_match_numpy_loops(oracle_pm):
    from pmesh import pm as opm
    if opm._load_cic() is None:
        pytest.skip("oracle/_build/libcic_oracle.so not built")
    rng = numpy.random.default_rng(4)
    for Nmesh, Box in [([8, 8, 8], 8.0), ([5, 7, 6], [3.0, 9.0, 4.0]), ([8, 8], 4.0)]:
        pm = oracle_pm(Nmesh, Box)
        nd = len(Nmesh)
        pos = rng.uniform(-1, 1, size=(400, nd)) * numpy.asarray(Box) * 2
        m = rng.uniform(0.5, 2, size=400)
        f = rng.standard_normal(pm.rshape)
        v = rng.standard_normal((400, nd))
        res = {}
        for use in (True, False):
            opm.use_c(use)
            res[use] = [pm.paint(pos, m).value, pm.paint(pos, 1.0).value,
                        opm._cic_paint_gradient(pm, pos, m, v), opm._cic_readout(pm, f, pos),
                        opm._cic_readout_gradient(pm, f, pos), pm.cell_keys(pos),
                        pm.decompose(pos).gather(v)]
        opm.use_c(True)
        for a, b in zip(res[True], res[False]):
            assert numpy.allclose(a, b, rtol=1e-13, atol=1e-13)
        assert numpy.array_equal(res[True][5], res[False][5])


def test_background_against_product():
    """oracle (odeint) and product (DOP853) growth factors agree far below the 1e-10 bar"""
    import oracle
    oracle.activate(shims=False)
    from fastpm.background import MatterDominated as O
    from vmad_b200.background import MatterDominated as P
    st = numpy.linspace(0.1, 1, 11)
    sup = numpy.sort(numpy.concatenate([(st[1:] * st[:-1]) ** 0.5, st]))
    o, p = O(0.3075, a=sup), P(0.3075, a=sup)
    for fn in ['E', 'D1', 'D2', 'f1', 'f2', 'Gp', 'gp', 'Gf', 'gf']:
        for a in sup:
            vo, vp = getattr(o, fn)(a), getattr(p, fn)(a)
            assert abs(vo - vp) < 1e-11 * abs(vo), (fn, a, vo, vp)
    assert abs(p.D1(1.0) - 1.0) < 1e-14
    # matter-dominated limit (dense default support): D2 ~ -3/7 D1^2, f1 ~ 1
    pd = P(0.3075)
    assert abs(pd.f1(0.01) - 1) < 1e-2
    assert abs(pd.D2(0.01) / pd.D1(0.01) ** 2 + 3. / 7) < 1e-2


@pytest.mark.parametrize("P", [2, 4, 8])
def test_p_rank_routing_oracle_matches_one_rank(oracle_pm, P):
    """the P-rank restatement of decompose (oracle/pmesh/domain.py) pinned by what the reference
    pins for layouts (test_fastpm.py:151-206: exchange then paint == paint): painting each rank's
    arrivals onto its own slab only and stacking the slabs equals the 1-rank paint; every particle
    reaches exactly the owners of the two planes it touches, in input order per destination"""
    from pmesh import domain
    N, L = 16, 100.0
    pm = oracle_pm([N] * 3, L)
    rng = numpy.random.default_rng(40 + P)
    pos = [rng.uniform(-0.4 * L, 1.4 * L, size=(700 + 13 * r, 3)) for r in range(P)]
    for r in range(P):                      # plane edges, the box edges, -eps
        pos[r][:6, 0] = [0.0, L, -1e-14, L * (1 - 1e-16), L / N * (N // P), L / N * (N // P) - 1e-13]
    mass = [rng.uniform(0.5, 1.5, size=len(p)) for p in pos]
    d = domain.decompose_ranks(pos, [N] * 3, L)
    ppr = N // P
    full = numpy.zeros((N, N, N))
    for dst in range(P):
        xs = numpy.concatenate([pos[s][d['send_idx'][s][d['send_start'][s][dst]:d['send_start'][s][dst + 1]]]
                                for s in range(P)])
        ms = numpy.concatenate([mass[s][d['send_idx'][s][d['send_start'][s][dst]:d['send_start'][s][dst + 1]]]
                                for s in range(P)])
        assert numpy.array_equal(xs, numpy.concatenate([pos[s] for s in range(P)], axis=0)[
            numpy.concatenate([numpy.cumsum([0] + [len(p) for p in pos])[s] + d['recv_idx'][dst][d['recv_src'][dst] == s]
                               for s in range(P)])])
        full[dst * ppr:(dst + 1) * ppr] = pm.paint(xs, ms).value[dst * ppr:(dst + 1) * ppr]
        # the local cell sort of what arrived is non-decreasing in the ghost-plane keys and stable
        keys = domain.ghost_keys(xs, [N] * 3, L, P, dst)
        assert (numpy.diff(keys[d['perm'][dst]]) >= 0).all()
        assert numpy.array_equal(d['perm'][dst], numpy.argsort(keys, kind='stable'))
    one = pm.paint(numpy.concatenate(pos), numpy.concatenate(mass)).value
    assert numpy.abs(full - one).max() < 1e-12
    for s in range(P):
        i0 = domain.floor_plane(pos[s][:, 0], N, L)
        copies = numpy.bincount(d['send_idx'][s], minlength=len(pos[s]))
        want = 1 + ((i0 // ppr) != (numpy.mod(i0 + 1, N) // ppr))
        assert numpy.array_equal(copies, want)
        for dst in range(P):                # input order within a destination
            seg = d['send_idx'][s][d['send_start'][s][dst]:d['send_start'][s][dst + 1]]
            assert (numpy.diff(seg) > 0).all()
    assert d['counts'].sum() == sum(len(x) for x in d['send_idx'])
