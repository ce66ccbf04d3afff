"""Parity of every CUDA kernel (through the C ABI / device pm) with the numpy oracle.

Bars (BASELINE.json north_star): bit-exact for cell keys, sort permutations and cell offsets;
fp64 fields and particle values within 1e-12 relative L2.
"""
import numpy
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-12


def edge_positions(rng, Nmesh, BoxSize, n):
    """random positions plus every nasty case: 0, L, -eps, L-eps, exact cell edges, far outside"""
    nd = len(Nmesh)
    L = numpy.empty(nd)
    L[...] = BoxSize
    x = rng.uniform(0, 1, size=(n, nd)) * L
    cell = L / numpy.array(Nmesh)
    eps = numpy.finfo('f8').eps
    special = []
    for d in range(nd):
        for v in [0.0, L[d], -eps * L[d], L[d] * (1 - eps), -L[d] * 0.3, L[d] * 1.7, -5.25 * L[d],
                  7.5 * L[d], cell[d], 3 * cell[d], cell[d] * (1 - eps), cell[d] * (1 + eps)]:
            p = rng.uniform(0, 1, size=nd) * L
            p[d] = v
            special.append(p)
    lattice = (numpy.indices(Nmesh).reshape(nd, -1).T * cell)[:min(n, 4096)]
    return numpy.concatenate([x, numpy.array(special), lattice], axis=0)


CASES = [([8, 8, 8], 16.0, 500), ([32, 32, 32], 100.0, 40000), ([20, 12, 36], [7.0, 3.0, 11.0], 9000),
         ([64, 64, 64], 200.0, 1 << 20)]


@pytest.mark.parametrize("Nmesh,BoxSize,n", CASES)
def test_keys_and_layout_bit_exact(oracle_pm, device_pm, Nmesh, BoxSize, n):
    rng = numpy.random.default_rng(1)
    opm = oracle_pm(Nmesh, BoxSize)
    dpm = device_pm(Nmesh, BoxSize)
    x = edge_positions(rng, Nmesh, BoxSize, n)
    okeys = opm.cell_keys(x)
    dkeys = dpm.cell_keys(x).cpu().numpy()
    assert numpy.array_equal(okeys, dkeys)
    olay = opm.decompose(x)
    dlay = dpm.decompose(x, keep_index=True)
    assert numpy.array_equal(dlay.keys.cpu().numpy(), okeys)
    assert numpy.array_equal(dlay.indices.cpu().numpy(), olay.indices)
    cs = dlay.cell_start.cpu().numpy()
    ref = numpy.concatenate([[0], numpy.cumsum(numpy.bincount(okeys, minlength=opm.Ntot))])
    assert numpy.array_equal(cs[:opm.Ntot + 1], ref)
    assert cs[-1] == len(x)
    # exchange / gather routing
    ex = dlay.exchange(x).numpy()
    assert numpy.array_equal(ex, olay.exchange(x))
    v = rng.standard_normal(len(x))
    assert numpy.array_equal(dlay.gather(dlay.exchange(v)).numpy(), v)
    assert numpy.array_equal(dlay.gather(v).numpy(), olay.gather(v))


def test_layout_crowded_cells(oracle_pm, device_pm):
    """many particles per cell (beyond the insertion-sort limit) keep the stable order"""
    rng = numpy.random.default_rng(2)
    opm = oracle_pm([4, 4, 4], 4.0)
    dpm = device_pm([4, 4, 4], 4.0)
    x = rng.uniform(0, 4, size=(50000, 3))
    x[:20000] = [1.5, 2.5, 3.5]            # 20000 particles in one cell
    x[20000:20100] = [0.5, 0.5, 0.5]
    olay = opm.decompose(x)
    dlay = dpm.decompose(x)
    assert numpy.array_equal(dlay.indices.cpu().numpy(), olay.indices)


@pytest.fixture(params=['aggregated', 'tile', 'gather'])
def paint_algorithm(request):
    from vmad_b200 import pm as dpm_mod
    dpm_mod.set_paint_algorithm(request.param)
    yield request.param
    dpm_mod.set_paint_algorithm('aggregated')


@pytest.mark.parametrize("Nmesh,BoxSize,n", CASES)
def test_paint_parity(oracle_pm, device_pm, paint_algorithm, Nmesh, BoxSize, n):
    rng = numpy.random.default_rng(3)
    opm = oracle_pm(Nmesh, BoxSize)
    dpm = device_pm(Nmesh, BoxSize)
    x = edge_positions(rng, Nmesh, BoxSize, n)
    m = rng.uniform(0.5, 1.5, size=len(x))
    ref = opm.paint(x, 1.0).value
    got = dpm.paint(x, 1.0).numpy()
    assert rel_l2(got, ref) < TOL
    assert abs(got.sum() - len(x)) < 1e-9 * len(x)
    ref = opm.paint(x, m).value
    assert rel_l2(dpm.paint(x, m).numpy(), ref) < TOL
    # with an explicit layout, and through exchange (tagged, no re-sort)
    lay = dpm.decompose(x)
    assert rel_l2(dpm.paint(x, m, layout=lay).numpy(), ref) < TOL
    assert rel_l2(dpm.paint(lay.exchange(x), lay.exchange(m)).numpy(), ref) < TOL
    # atomic baseline kernel and the row-gather ablation kernel
    assert rel_l2(dpm.paint_atomic(x, m).numpy(), ref) < TOL
    assert rel_l2(dpm.paint_rows(x, m).numpy(), ref) < TOL
    if paint_algorithm in ('tile', 'gather'):
        # fixed summation order: run-to-run bit-identical
        a = dpm.paint(x, m).numpy()
        b = dpm.paint(x, m).numpy()
        assert numpy.array_equal(a, b)
    # three columns in one pass (reverse sweep of the force operator)
    if len(Nmesh) == 3:
        m3 = rng.standard_normal((len(x), 3))
        xs, ms = lay.exchange(x), lay.exchange(m3)
        got3 = dpm.paint_cols3(xs, ms)
        for d in range(3):
            assert rel_l2(got3[d].numpy(), opm.paint(x, m3[:, d]).value) < TOL


def test_paint_uniform_grid_is_one(device_pm):
    """test_fastpm.py:158-159,176-177: one particle per cell paints exactly 1 per cell"""
    for shift in (0.0, 0.1, 0.5):
        dpm = device_pm([16, 16, 16], 50.0)
        q = dpm.generate_uniform_particle_grid(shift=shift)
        mesh = dpm.paint(q, 1.0).numpy()
        assert numpy.abs(mesh - 1.0).max() < 1e-13


@pytest.mark.parametrize("Nmesh,BoxSize,n", CASES[:3])
def test_readout_family_parity(oracle_pm, device_pm, paint_algorithm, Nmesh, BoxSize, n):
    rng = numpy.random.default_rng(4)
    opm = oracle_pm(Nmesh, BoxSize)
    dpm = device_pm(Nmesh, BoxSize)
    x = edge_positions(rng, Nmesh, BoxSize, n)
    f = rng.standard_normal(Nmesh)
    fdot = rng.standard_normal(Nmesh)
    v = rng.standard_normal(len(x))
    xdot = rng.standard_normal(x.shape)
    m = rng.uniform(0.5, 1.5, size=len(x))
    of = opm.create('real', value=f)
    df = dpm.create('real', value=f)
    assert rel_l2(df.readout(x).numpy(), of.readout(x)) < TOL
    lay = dpm.decompose(x)
    assert rel_l2(df.readout(x, layout=lay).numpy(), of.readout(x)) < TOL
    # readout_vjp
    omesh, opos = of.readout_vjp(x, v)
    dmesh, dpos = df.readout_vjp(x, v)
    assert rel_l2(dmesh.numpy(), omesh.value) < TOL
    assert rel_l2(dpos.numpy(), opos) < TOL
    dmesh, dpos = df.readout_vjp(x, v, layout=lay)
    assert rel_l2(dmesh.numpy(), omesh.value) < TOL
    assert rel_l2(dpos.numpy(), opos) < TOL
    # readout_jvp (both terms, each alone)
    ref = of.readout_jvp(x, v_self=opm.create('real', value=fdot), v_pos=xdot)
    assert rel_l2(df.readout_jvp(x, v_self=fdot, v_pos=xdot).numpy(), ref) < TOL
    assert rel_l2(df.readout_jvp(x, v_pos=xdot).numpy(), of.readout_jvp(x, v_pos=xdot)) < TOL
    assert rel_l2(df.readout_jvp(x, v_self=fdot).numpy(),
                  of.readout_jvp(x, v_self=opm.create('real', value=fdot))) < TOL
    # paint_vjp
    opos, omass = opm.paint_vjp(of, x, mass=m)
    dpos, dmass = dpm.paint_vjp(df, x, mass=m)
    assert rel_l2(dpos.numpy(), opos) < TOL
    assert rel_l2(dmass.numpy(), omass) < TOL
    opos, omass = opm.paint_vjp(of, x, mass=1.0)
    dpos, dmass = dpm.paint_vjp(df, x, mass=1.0, layout=lay)
    assert rel_l2(dpos.numpy(), opos) < TOL
    assert rel_l2(dmass.numpy(), omass) < TOL
    # paint_jvp
    ref = opm.paint_jvp(x, v_mass=v, mass=m, v_pos=xdot).value
    assert rel_l2(dpm.paint_jvp(x, v_mass=v, mass=m, v_pos=xdot).numpy(), ref) < TOL
    ref = opm.paint_jvp(x, mass=1.0, v_pos=xdot).value
    assert rel_l2(dpm.paint_jvp(x, mass=1.0, v_pos=xdot).numpy(), ref) < TOL
    ref = opm.paint_jvp(x, v_mass=v, mass=m).value
    assert rel_l2(dpm.paint_jvp(x, v_mass=v, mass=m).numpy(), ref) < TOL
    # fused three-field readout
    fs = [rng.standard_normal(Nmesh) for _ in range(3)]
    got = dpm.readout3(*[dpm.create('real', value=a) for a in fs], x).numpy()
    ref = numpy.stack([opm.create('real', value=a).readout(x) for a in fs], axis=-1)
    assert rel_l2(got, ref) < TOL


@pytest.mark.parametrize("Nmesh", [[4, 4], [8, 8], [4, 4, 4], [32, 32, 32], [20, 12, 36], [128, 128, 128]])
def test_fft_parity(oracle_pm, device_pm, Nmesh):
    rng = numpy.random.default_rng(5)
    opm = oracle_pm(Nmesh, 8.0)
    dpm = device_pm(Nmesh, 8.0)
    r = rng.standard_normal(Nmesh)
    c = rng.standard_normal(opm.cshape) + 1j * rng.standard_normal(opm.cshape)  # NOT hermitian-consistent
    orf, dr = opm.create('real', value=r), dpm.create('real', value=r)
    oc, dc = opm.create('complex', value=c), dpm.create('complex', value=c)
    assert rel_l2(dr.r2c().numpy(), orf.r2c().value) < TOL
    assert rel_l2(dr.c2r_vjp().numpy(), orf.c2r_vjp().value) < TOL
    assert rel_l2(dc.c2r().numpy(), oc.c2r().value) < TOL
    assert rel_l2(dc.r2c_vjp().numpy(), oc.r2c_vjp().value) < TOL
    assert numpy.array_equal(dc.numpy(), c)  # c2r must not clobber its input
    # round trip identity (test_fastpm.py:25-37)
    assert rel_l2(dr.r2c().c2r().numpy(), r) < TOL
    # reductions
    assert abs(dr.cnorm() - orf.cnorm()) < 1e-12 * abs(orf.cnorm())
    assert abs(dc.cnorm() - oc.cnorm()) < 1e-12 * abs(oc.cnorm())
    c2 = rng.standard_normal(opm.cshape) + 1j * rng.standard_normal(opm.cshape)
    assert abs(dc.cdot(c2) - oc.cdot(c2)) < 1e-12 * abs(oc.cdot(c2))
    # Parseval (test_fastpm.py:142)
    r2 = rng.standard_normal(Nmesh)
    lhs = dr.r2c().cdot(dpm.create('real', value=r2).r2c()).real
    assert abs(lhs - (r * r2).sum() / opm.Ntot) < 1e-12 * abs(lhs) + 1e-15


@pytest.mark.parametrize("Nmesh", [[8, 8], [8, 8, 8], [20, 12, 36]])
def test_transfer_parity(oracle_pm, device_pm, Nmesh):
    rng = numpy.random.default_rng(6)
    BoxSize = 13.0
    opm = oracle_pm(Nmesh, BoxSize)
    dpm = device_pm(Nmesh, BoxSize)
    c = rng.standard_normal(opm.cshape) + 1j * rng.standard_normal(opm.cshape)
    oc, dc = opm.create('complex', value=c), dpm.create('complex', value=c)

    def laplace(k):
        k2 = k.normp(2)
        bad = k2 == 0
        k2[bad] = 1
        k2 = -1 / k2
        k2[bad] = 0
        return k2

    def grad(d):
        def f(k):
            h = BoxSize / Nmesh[d]
            w = k[d] * h
            return -1j * (1 / (6.0 * h) * (8 * numpy.sin(w) - numpy.sin(2 * w)))
        return f

    def mixed(k):
        return (1.0 + sum(ki ** 2 for ki in k)) ** -0.5 + 0.25j * k[0]

    for tf in [laplace, grad(0), grad(len(Nmesh) - 1), mixed]:
        ref = oc.apply(lambda k, v: v * tf(k)).value
        got = dc.apply(lambda k, v: v * tf(k)).numpy()
        assert rel_l2(got, ref) < TOL
        ref = oc.apply(lambda k, v: v * numpy.conj(tf(k))).value
        got = dc.apply(lambda k, v: v * numpy.conj(tf(k))).numpy()
        assert rel_l2(got, ref) < TOL
    # the fused Green's function kernel reproduces numpy's -1/k^2 to the bit
    got = dpm._transfer(dc, laplace=True).numpy()
    ref = oc.apply(lambda k, v: v * laplace(k)).value
    assert numpy.array_equal(got, ref)
    for kind in ('circular', 'index'):
        ref = oc.apply(lambda k, v: v * (1 + k[0] + 2 * k[-1]), kind=kind).value
        got = dc.apply(lambda k, v: v * (1 + k[0] + 2 * k[-1]), kind=kind).numpy()
        assert rel_l2(got, ref) < TOL


def test_elementwise_and_numpy_protocol(device_pm):
    from vmad_b200.pm import DeviceArray
    rng = numpy.random.default_rng(7)
    a = rng.standard_normal((1000, 3))
    b = rng.standard_normal((1000, 3))
    w = rng.standard_normal(1000)
    A, B, W = DeviceArray.from_host(a), DeviceArray.from_host(b), DeviceArray.from_host(w)
    assert numpy.array_equal((A + B).numpy(), a + b)
    assert numpy.array_equal((A - B).numpy(), a - b)
    assert numpy.array_equal((A * B).numpy(), a * b)
    assert numpy.array_equal((A / B).numpy(), a / b)
    assert numpy.array_equal((A * 2.5).numpy(), a * 2.5)
    assert numpy.array_equal((2.5 * A).numpy(), a * 2.5)
    assert numpy.array_equal((numpy.float64(2.5) * A).numpy(), a * 2.5)
    assert numpy.array_equal((-A).numpy(), -a)
    assert (A + 0) is A and (0 + A) is A
    assert numpy.array_equal((A + 1.5).numpy(), a + 1.5)
    assert numpy.array_equal((1.5 - A).numpy(), 1.5 - a)
    assert numpy.array_equal((A * W).numpy(), a * w[:, None])
    assert numpy.allclose((A ** 2).numpy(), a ** 2, rtol=0, atol=0)
    assert abs(A.sum() - a.sum()) < 1e-10
    assert abs(numpy.sum(A * B) - (a * b).sum()) < 1e-10
    cols = [DeviceArray.from_host(a[:, i].copy()) for i in range(3)]
    S = numpy.stack(cols, axis=-1)
    assert isinstance(S, DeviceArray) and numpy.array_equal(S.numpy(), a)
    T = numpy.take(A, 1, axis=-1)
    assert isinstance(T, DeviceArray) and numpy.array_equal(T.numpy(), a[:, 1])
    assert numpy.array_equal((A + b).numpy(), a + b)  # host operand is uploaded


def test_paint_readout_adjoint_large(device_pm):
    """size-independent property at a BASELINE-size problem: <paint(x, m), g> == <m, readout(g, x)>"""
    from vmad_b200.pm import DeviceArray
    import torch
    N = 256
    dpm = device_pm([N, N, N], 800.0)
    g = torch.Generator(device='cuda').manual_seed(11)
    n = 1 << 24
    x = DeviceArray(torch.rand((n, 3), generator=g, device='cuda', dtype=torch.float64) * 800.0)
    m = DeviceArray(torch.rand((n,), generator=g, device='cuda', dtype=torch.float64) + 0.5)
    f = dpm.create('real')
    f.t.copy_(torch.randn((N, N, N), generator=g, device='cuda', dtype=torch.float64))
    lhs = dpm.paint(x, m).dot(f)
    rhs = f.readout(x).dot(m)
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)
    total = dpm.paint(x, m).sum()
    assert abs(total - m.sum()) < 1e-11 * total


def test_edge_cases_empty_single_and_odd_meshes(oracle_pm, device_pm, paint_algorithm):
    """empty and one-particle inputs, odd mesh sizes (no Nyquist mode), a mesh axis of length 1"""
    rng = numpy.random.default_rng(8)
    for Nmesh, Box in [([5, 7, 9], [5.0, 3.5, 18.0]), ([1, 6, 4], 6.0), ([3, 1, 5], 2.0), ([9, 9], 4.5)]:
        opm = oracle_pm(Nmesh, Box)
        dpm = device_pm(Nmesh, Box)
        nd = len(Nmesh)
        L = numpy.empty(nd)
        L[...] = Box
        # empty particle set: a zero mesh, empty readouts, an empty layout
        empty = numpy.zeros((0, nd))
        assert numpy.array_equal(dpm.paint(empty, 1.0).numpy(), numpy.zeros(Nmesh))
        f = rng.standard_normal(Nmesh)
        df, of = dpm.create('real', value=f), opm.create('real', value=f)
        assert df.readout(empty).numpy().shape == (0,)
        lay = dpm.decompose(empty)
        assert lay.exchange(empty).numpy().shape == (0, nd)
        assert lay.gather(numpy.zeros((0,))).numpy().shape == (0,)
        # one particle, and a handful, on this odd / degenerate mesh
        for n in (1, 37):
            x = rng.uniform(-1, 2, size=(n, nd)) * L
            m = rng.uniform(0.5, 1.5, size=n)
            assert rel_l2(dpm.paint(x, m).numpy(), opm.paint(x, m).value) < TOL
            assert rel_l2(df.readout(x).numpy(), of.readout(x)) < TOL
            dmesh, dpos = df.readout_vjp(x, m)
            omesh, opos = of.readout_vjp(x, m)
            assert rel_l2(dmesh.numpy(), omesh.value) < TOL
            assert rel_l2(dpos.numpy(), opos) < TOL
            assert numpy.array_equal(dpm.decompose(x).indices.cpu().numpy(), opm.decompose(x).indices)
        # transforms and a k-space filter on odd sizes
        c = rng.standard_normal(opm.cshape) + 1j * rng.standard_normal(opm.cshape)
        dc, oc = dpm.create('complex', value=c), opm.create('complex', value=c)
        assert rel_l2(df.r2c().numpy(), of.r2c().value) < TOL
        assert rel_l2(dc.c2r().numpy(), oc.c2r().value) < TOL
        assert abs(dc.cnorm() - oc.cnorm()) < 1e-12 * abs(oc.cnorm())
        tf = lambda k: 1.0 / (1.0 + sum(ki ** 2 for ki in k)) + 0.5j * k[-1]
        assert rel_l2(dc.apply(lambda k, v: v * tf(k)).numpy(), oc.apply(lambda k, v: v * tf(k)).value) < TOL


def test_routing_helpers_bit_exact(device_pm):
    """the index plumbing of the cross-rank layout, driven on one GPU through the C ABI:
    vpm_route_compose, vpm_take_rows2, masked vpm_scatter_add_rows, and the fused gather of
    vpm_readout3 (out_row) -- all bit-exact against numpy indexing"""
    import ctypes
    import torch
    from vmad_b200 import _capi as C
    from vmad_b200.pm import runtime
    rt = runtime()
    rng = numpy.random.default_rng(21)
    nsend, nrecv = 5000, 5200
    send_idx = rng.integers(0, 4000, size=nsend).astype('i4')
    perm = rng.permutation(nrecv).astype('i4')                # sorted slot -> receive row
    recv_lo, send_lo, nself = 700, 1200, 3100                 # self segment in both orders
    dev = lambda a: torch.from_numpy(numpy.ascontiguousarray(a)).cuda()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    take = torch.empty(nrecv, dtype=torch.int32, device='cuda')
    remote = torch.empty(nsend, dtype=torch.int32, device='cuda')
    d_perm, d_send = dev(perm), dev(send_idx)
    inv = torch.empty(nrecv - nself, dtype=torch.int32, device='cuda')
    C.check(C.lib.vpm_route_compose(rt.ctx, P(d_perm), nrecv, P(d_send), nsend, recv_lo, recv_lo + nself,
                                    send_lo, P(take), P(remote), P(inv)))
    want_inv = numpy.empty(nrecv, dtype='i4')
    want_inv[perm] = numpy.arange(nrecv, dtype='i4')
    want_inv = numpy.concatenate([want_inv[:recv_lo], want_inv[recv_lo + nself:]])
    assert numpy.array_equal(inv.cpu().numpy(), want_inv)
    self_row = (perm >= recv_lo) & (perm < recv_lo + nself)
    want_take = numpy.where(self_row, send_idx[numpy.clip(send_lo + perm - recv_lo, 0, nsend - 1)], -1 - perm)
    want_remote = send_idx.copy()
    want_remote[send_lo:send_lo + nself] = -1
    assert numpy.array_equal(take.cpu().numpy(), want_take)
    assert numpy.array_equal(remote.cpu().numpy(), want_remote)
    # take_rows2: home rows where take >= 0, receive-area rows elsewhere
    home = rng.standard_normal((4000, 3))
    recv = rng.standard_normal((nrecv, 3))
    out = torch.empty((nrecv, 3), dtype=torch.float64, device='cuda')
    d_home, d_recv = dev(home), dev(recv)
    C.check(C.lib.vpm_take_rows2(rt.ctx, P(d_home), P(d_recv), P(take), nrecv, 3, P(out)))
    want = numpy.where(self_row[:, None], home[numpy.clip(want_take, 0, None)], recv[perm])
    assert numpy.array_equal(out.cpu().numpy(), want)
    # masked scatter-add: negative rows are skipped
    acc = torch.zeros((4000, 3), dtype=torch.float64, device='cuda')
    vals = rng.standard_normal((nsend, 3))
    d_vals = dev(vals)
    C.check(C.lib.vpm_scatter_add_rows(rt.ctx, P(d_vals), P(remote), nsend, 3, P(acc)))
    ref = numpy.zeros((4000, 3))
    keep = want_remote >= 0
    numpy.add.at(ref, want_remote[keep], vals[keep])
    assert rel_l2(acc.cpu().numpy(), ref) < 1e-15
    # readout3 with the layout's permutation as out_row == readout3 followed by Layout.gather
    dpm = device_pm([16, 16, 16], 50.0)
    x = rng.uniform(0, 50.0, size=(3000, 3))
    fs = [dpm.create('real', value=rng.standard_normal((16, 16, 16))) for _ in range(3)]
    lay = dpm.decompose(x)
    xs = lay.exchange(x)
    a = dpm.readout3(fs[0], fs[1], fs[2], xs, gather=lay).numpy()
    b = lay.gather(dpm.readout3(fs[0], fs[1], fs[2], xs)).numpy()
    assert numpy.array_equal(a, b)


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("Nmesh,BoxSize,n", [([8, 8, 8], 16.0, 3000), ([32, 32, 32], 100.0, 100000),
                                             ([64, 16, 24], [200.0, 30.0, 11.0], 50000)])
def test_route_build_bit_exact_vs_p_rank_oracle(oracle_pm, device_pm, P, Nmesh, BoxSize, n):
    """`north_star`: bit-exact exchange routing.  vpm_route_build (owner of the floor plane + ghost
    copy to the owner of plane + 1, per-destination input order, per-destination counts) against
    the P-rank restatement of pmesh's decompose (oracle/pmesh/domain.py), for the particles of one
    rank -- the routing of a rank does not depend on the others -- including every edge position"""
    import ctypes
    import torch
    from pmesh import domain
    from vmad_b200 import _capi as C
    rng = numpy.random.default_rng(5 + P)
    dpm = device_pm(Nmesh, BoxSize)
    x = edge_positions(rng, Nmesh, BoxSize, n)
    L0 = numpy.broadcast_to(numpy.asarray(BoxSize, dtype='f8'), (3,))[0]
    # particles exactly on the slab boundaries and one ulp below them
    ppr = Nmesh[0] // P
    edges = numpy.array([k * ppr * (L0 / Nmesh[0]) for k in range(P + 1)])
    extra = rng.uniform(0, 1, size=(2 * len(edges), 3)) * numpy.asarray(BoxSize)
    extra[:len(edges), 0] = edges
    extra[len(edges):, 0] = numpy.nextafter(edges, -numpy.inf)
    x = numpy.concatenate([x, extra])
    want_idx, want_start = domain.route(x, Nmesh, BoxSize, P)
    xd = torch.from_numpy(numpy.ascontiguousarray(x)).cuda()
    send_idx = torch.empty(2 * len(x), dtype=torch.int32, device='cuda')
    start = (ctypes.c_int32 * (P + 1))()
    rt = dpm.rt
    rt.sync_stream()
    C.check(C.lib.vpm_route_build(rt.ctx, ctypes.byref(dpm._mesh), P, ctypes.c_void_p(xd.data_ptr()), len(x),
                                  ctypes.c_void_p(send_idx.data_ptr()), send_idx.shape[0], start))
    got_start = numpy.array(list(start), dtype='i8')
    assert numpy.array_equal(got_start, want_start)
    assert numpy.array_equal(send_idx[:got_start[-1]].cpu().numpy(), want_idx)


@pytest.mark.parametrize("npart", [60000, 1 << 21])
@pytest.mark.parametrize("P", [2, 4, 8])
def test_route_build_static_bit_exact(oracle_pm, device_pm, P, npart):
    """the fixed-capacity variant of the routing (no host synchronisation): the same lists inside
    fixed segments, -1 in the unused tails, device-side counts, overflow flag only when a segment
    is too small"""
    import ctypes
    import torch
    from pmesh import domain
    from vmad_b200 import _capi as C
    Nmesh, BoxSize = [32, 32, 32], 100.0
    rng = numpy.random.default_rng(50 + P)
    dpm = device_pm(Nmesh, BoxSize)
    x = edge_positions(rng, Nmesh, BoxSize, npart)
    want_idx, want_start = domain.route(x, Nmesh, BoxSize, P)
    want_cnt = numpy.diff(want_start)
    xd = torch.from_numpy(numpy.ascontiguousarray(x)).cuda()
    rt = dpm.rt
    rt.sync_stream()
    for slack, expect_overflow in ((1.0, False), (1.5, False), (0.5, True)):
        cap = numpy.maximum((want_cnt * slack).astype('i8') + (5 if slack > 1 else 0), 1)
        seg = numpy.concatenate([[0], numpy.cumsum(cap)])
        send_idx = torch.empty(int(seg[-1]), dtype=torch.int32, device='cuda')
        counts = torch.zeros(P + 1, dtype=torch.int32, device='cuda')
        cseg = (ctypes.c_int32 * (P + 1))(*[int(v) for v in seg])
        C.check(C.lib.vpm_route_build_static(rt.ctx, ctypes.byref(dpm._mesh), P, ctypes.c_void_p(xd.data_ptr()),
                                             len(x), cseg, ctypes.c_void_p(send_idx.data_ptr()),
                                             ctypes.c_void_p(counts.data_ptr())))
        got = send_idx.cpu().numpy()
        cnt = counts.cpu().numpy()
        assert bool(cnt[P]) == expect_overflow
        assert numpy.array_equal(cnt[:P], numpy.minimum(want_cnt, cap))
        for d in range(P):
            w = want_idx[want_start[d]:want_start[d + 1]][:cap[d]]
            s = got[seg[d]:seg[d + 1]]
            assert numpy.array_equal(s[:len(w)], w)
            assert (s[len(w):] == -1).all()


@pytest.mark.parametrize("shape,pad", [((16, 12, 32), 0), ((8, 20, 36), 0), ((6, 8, 8), 2), ((5, 7, 10), 0),
                                       ((64, 64, 64), 0), ((9, 16, 520), 2), ((40, 30, 256), 0), ((33, 13, 132), 2)])
def test_fd4_stencil_and_adjoint(device_pm, shape, pad):
    """4-point finite-difference gradient F_d = -D_d phi (the real-space form of
    fastpm.py:316-323, order 1) and its adjoint against numpy rolls; the four-cells-per-thread
    kernels (n2 % 4 == 0, aligned arrays) must agree bit for bit with the scalar ones"""
    import ctypes
    import torch
    from vmad_b200 import _capi as C
    from vmad_b200.pm import runtime, _ptr
    rt = runtime()
    rng = numpy.random.default_rng(11)
    n0, n1, n2 = shape
    h = numpy.array([0.7, 1.3, 0.9])
    phi = rng.standard_normal((n0 + 2 * pad, n1, n2))
    Fb = rng.standard_normal((3, n0 + 2 * pad, n1, n2))

    def D(f, axis):
        # periodic along axes 1, 2; along axis 0 periodic when pad == 0, else the halo planes are given
        r = lambda k: numpy.roll(f, -k, axis=axis)
        d = (8.0 * (r(1) - r(-1)) - (r(2) - r(-2))) * (1.0 / (12.0 * h[axis]))
        return d[pad:pad + n0] if pad else d

    want = [-D(phi, d) for d in range(3)]
    want_adj = D(Fb[2], 2) + D(Fb[1], 1) + D(Fb[0], 0)
    n3 = (ctypes.c_int64 * 3)(n0, n1, n2)
    h3 = (ctypes.c_double * 3)(*h)

    def run(misalign):
        # one spare element in front: misalign = 1 shifts every array off 16-byte alignment -> scalar kernels
        def dev(a):
            t = torch.empty(a.size + 2, dtype=torch.float64, device=rt.device)
            v = t[misalign:misalign + a.size].view(a.shape)
            v.copy_(torch.from_numpy(numpy.ascontiguousarray(a)))
            return t, v
        keep = []
        _, dphi = dev(phi); keep.append(_)
        outs = []
        for _ in range(3):
            t, v = dev(numpy.zeros((n0, n1, n2))); keep.append(t); outs.append(v)
        rt.sync_stream()
        C.check(C.lib.vpm_fd4_neg_gradient_slab(rt.ctx, n3, h3, pad, _ptr(dphi), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])))
        fb = []
        for d in range(3):
            t, v = dev(Fb[d]); keep.append(t); fb.append(v)
        t, pb = dev(numpy.zeros((n0, n1, n2))); keep.append(t)
        C.check(C.lib.vpm_fd4_neg_gradient_adjoint_slab(rt.ctx, n3, h3, pad, _ptr(fb[0]), _ptr(fb[1]), _ptr(fb[2]), _ptr(pb)))
        rt.synchronize()
        return [o.cpu().numpy() for o in outs], pb.cpu().numpy()

    got_s, got_adj_s = run(1)           # misaligned arrays: the scalar kernels
    try:
        for algo in (3, 1):             # marching (register window + shared-memory tile), four cells per thread
            C.check(C.lib.vpm_fd4_set_algorithm(algo))
            got, got_adj = run(0)
            for d in range(3):
                assert rel_l2(got[d], want[d]) < TOL
            assert rel_l2(got_adj, want_adj) < TOL
            for d in range(3):
                assert numpy.array_equal(got[d], got_s[d])
            assert numpy.array_equal(got_adj, got_adj_s)
    finally:
        C.check(C.lib.vpm_fd4_set_algorithm(2))


def test_readout_home_row_variants_vs_oracle(oracle_pm, device_pm):
    """the cross-rank forms of the force readout / its reverse sweep on ONE device:
    vpm_readout3_home and vpm_readout_grad4_home store `acc_in[h] + scale * value` at home rows
    h >= 0 and the slot-order copy at the slots whose home is elsewhere (h < 0), checked against
    the oracle readout / readout_vjp; also the fused one-rank epilogues (out_row, kick, base)"""
    import ctypes
    import torch
    from vmad_b200 import _capi as C
    from vmad_b200.pm import runtime, _ptr
    rng = numpy.random.default_rng(21)
    Nmesh, BoxSize, n = [16, 12, 20], [7.0, 3.0, 11.0], 5000
    opm = oracle_pm(Nmesh, BoxSize)
    dpm = device_pm(Nmesh, BoxSize)
    rt = runtime()
    x = edge_positions(rng, Nmesh, BoxSize, n)
    n = len(x)
    fs = [rng.standard_normal(Nmesh) for _ in range(4)]
    w3 = rng.standard_normal((n, 3))
    val = numpy.stack([opm.create('real', value=a).readout(x) for a in fs[:3]], axis=-1)
    # gradient of sum_d w3[:, d] readout(F_d, x) + readout(R, x) with respect to x
    grad = sum(opm.create('real', value=fs[d]).readout_vjp(x, w3[:, d])[1] for d in range(3)) \
        + opm.create('real', value=fs[3]).readout_vjp(x, numpy.ones(n))[1]
    # home rows: a random injection into a larger home array; a third of the slots live "elsewhere"
    nhome = n + 100
    home = rng.permutation(nhome)[:n].astype('i4')
    away = rng.uniform(size=n) < 0.33
    home_row = numpy.where(away, -1 - numpy.arange(n), home).astype('i4')
    home_row[::97] = numpy.iinfo('i4').min            # padding slots of a fixed-capacity layout
    away |= home_row == numpy.iinfo('i4').min
    acc_in = rng.standard_normal((nhome, 3))
    dev = lambda a, dt=torch.float64: torch.from_numpy(numpy.ascontiguousarray(a)).to(rt.device, dt)
    F = [dev(a) for a in fs]
    dx, dw, dh, dacc = dev(x), dev(w3), dev(home_row, torch.int32), dev(acc_in)
    mesh = ctypes.byref(dpm._mesh)
    rt.sync_stream()
    for scale, use_acc in ((0.7, True), (1.0, False)):
        slots = torch.full((n, 3), numpy.nan, dtype=torch.float64, device=rt.device)
        out = dacc.clone() if use_acc else torch.zeros((nhome, 3), dtype=torch.float64, device=rt.device)
        C.check(C.lib.vpm_readout3_home(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(dx), n, _ptr(dh),
                                        _ptr(slots), _ptr(dacc) if use_acc else None, scale, _ptr(out)))
        rt.synchronize()
        want = acc_in.copy() if use_acc else numpy.zeros((nhome, 3))
        want[home[~away]] += scale * val[~away]
        assert rel_l2(out.cpu().numpy(), want) < TOL
        got = slots.cpu().numpy()
        assert rel_l2(got[away], val[away]) < TOL
        assert numpy.isnan(got[~away]).all()            # slots whose home is here are not written
    slots = torch.full((n, 3), numpy.nan, dtype=torch.float64, device=rt.device)
    out = dacc.clone()
    C.check(C.lib.vpm_readout_grad4_home(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(F[3]), _ptr(dx),
                                         _ptr(dw), n, _ptr(dh), _ptr(slots), _ptr(dacc), _ptr(out)))
    rt.synchronize()
    want = acc_in.copy()
    want[home[~away]] += grad[~away]
    assert rel_l2(out.cpu().numpy(), want) < TOL
    assert rel_l2(slots.cpu().numpy()[away], grad[away]) < TOL
    # one-rank epilogues: rows stored through out_row, base added, y_out = y_in + fac * result
    perm = dev(home, torch.int32)
    g = torch.zeros((nhome, 3), dtype=torch.float64, device=rt.device)
    yo = torch.zeros((nhome, 3), dtype=torch.float64, device=rt.device)
    yin = rng.standard_normal((nhome, 3))
    C.check(C.lib.vpm_readout_grad4(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(F[3]), _ptr(dx), _ptr(dw),
                                    n, _ptr(perm), _ptr(dacc), _ptr(g), _ptr(dev(yin)), 0.3, _ptr(yo)))
    rt.synchronize()
    want = numpy.zeros((nhome, 3)); want[home] = grad + acc_in[home]
    wy = numpy.zeros((nhome, 3)); wy[home] = yin[home] + 0.3 * want[home]
    assert rel_l2(g.cpu().numpy(), want) < TOL
    assert rel_l2(yo.cpu().numpy(), wy) < TOL


def test_peer_put_strided_on_one_device():
    """vpm_peer_put_strided with one 'rank' (the receive area is a local buffer): rows land with
    the destination stride -- the layout the inverse slab transform reads in place"""
    import ctypes
    import torch
    from vmad_b200 import _capi as C
    from vmad_b200.pm import runtime
    rt = runtime()
    rt.sync_stream()
    rows, inner, s_row, d_row = 5, 24, 40, 64
    src = torch.arange(2 * rows * s_row, dtype=torch.float64, device=rt.device)      # complex pairs
    p = ctypes.c_void_p()
    h = ctypes.create_string_buffer(64)
    nbytes = 16 * (rows * d_row + 8)
    C.check(C.lib.vpm_peer_alloc(rt.ctx, nbytes, ctypes.byref(p), h))
    try:
        bufs = (ctypes.c_uint64 * 8)(p.value, 0, 0, 0, 0, 0, 0, 0)
        C.check(C.lib.vpm_peer_put_strided(rt.ctx, 1, rows, inner, s_row, 0, d_row, ctypes.c_void_p(src.data_ptr()),
                                           bufs, 3, None))
        out = torch.empty(nbytes // 8, dtype=torch.float64, device=rt.device)
        C.check(C.lib.vpm_axpby(rt.ctx, nbytes // 8, 1.0, p, 0.0, None, 0.0, ctypes.c_void_p(out.data_ptr())))
        rt.synchronize()
        got = out.cpu().numpy().reshape(-1, 2)
        want = numpy.zeros_like(got)
        s = src.cpu().numpy().reshape(-1, 2)
        for i in range(rows):
            want[3 + i * d_row: 3 + i * d_row + inner] = s[i * s_row: i * s_row + inner]
        assert numpy.array_equal(got, want)
    finally:
        C.check(C.lib.vpm_peer_free(rt.ctx, p))
