"""P-rank restatement of pmesh's particle routing (``ParticleMesh.decompose`` over a slab
decomposition; called at vmad/lib/fastpm.py:272) -- TEST INFRASTRUCTURE ONLY.

pmesh is an un-vendored, un-pinned dependency of the reference, so this
restates the published semantics of ``pmesh.domain.GridND.decompose`` with ``smoothing`` = half
the CIC support (1 cell) for a mesh sharded in slabs along axis 0: a particle is delivered to
every rank whose slab, widened by the smoothing, covers it -- for CIC the owner of its floor
plane ``i0`` and, when different, the owner of plane ``i0 + 1`` (the two planes it touches);
within one destination, particles keep their input order; what arrives on a rank is ordered by
source rank.  All ranks are simulated in one process with numpy integer arithmetic.

Used to check ``vpm_route_build`` (send lists, per-destination order, count matrix) and the
cross-rank ``DistLayout`` (receive order, local cell-sort permutation) bit for bit.
"""
import numpy


def floor_plane(pos0, nmesh0, box0):
    """wrapped floor plane of axis 0: one rounded multiply, floor, integer wrap (the oracle's
    cell arithmetic, oracle/pmesh/pm.py:cell_index)"""
    scale = numpy.float64(nmesh0) / numpy.float64(box0)
    u = numpy.asarray(pos0, dtype='f8') * scale
    return numpy.mod(numpy.floor(u).astype('i8'), nmesh0)


def route(pos, nmesh, boxsize, nranks):
    """routing of ONE rank's particles: (send_idx, send_start) with
    ``send_idx[send_start[d]:send_start[d + 1]]`` = indices bound for rank d in input order"""
    nmesh = numpy.asarray(nmesh, dtype='i8')
    box = numpy.empty(len(nmesh), dtype='f8')
    box[...] = boxsize
    n0 = int(nmesh[0])
    if n0 % nranks:
        raise ValueError("mesh planes must divide evenly over the ranks")
    ppr = n0 // nranks
    pos = numpy.asarray(pos, dtype='f8').reshape(-1, len(nmesh))
    i0 = floor_plane(pos[:, 0], n0, box[0])
    r0 = i0 // ppr
    r1 = numpy.mod(i0 + 1, n0) // ppr
    lists = []
    for d in range(nranks):
        lists.append(numpy.nonzero((r0 == d) | (r1 == d))[0])
    send_start = numpy.concatenate([[0], numpy.cumsum([len(l) for l in lists])]).astype('i8')
    send_idx = numpy.concatenate(lists).astype('i8') if lists else numpy.zeros(0, 'i8')
    return send_idx, send_start


def ghost_keys(pos, nmesh, boxsize, nranks, rank):
    """cell keys of particles that arrived on ``rank``: planes counted from the lower ghost
    plane (``sp = (i0 - start0 + 1) mod n0``), then rows, then columns"""
    nmesh = numpy.asarray(nmesh, dtype='i8')
    box = numpy.empty(len(nmesh), dtype='f8')
    box[...] = boxsize
    pos = numpy.asarray(pos, dtype='f8').reshape(-1, len(nmesh))
    scale = nmesh / box
    idx = numpy.mod(numpy.floor(pos * scale).astype('i8'), nmesh)
    ppr = int(nmesh[0]) // nranks
    sp = numpy.mod(idx[:, 0] - rank * ppr + 1, nmesh[0])
    if (sp > ppr).any():
        raise ValueError("a particle was routed to a rank that does not hold its planes")
    return (sp * nmesh[1] + idx[:, 1]) * nmesh[2] + idx[:, 2]


def decompose_ranks(pos_per_rank, nmesh, boxsize):
    """all ranks at once.  Returns a dict:
    ``send_idx[s]``, ``send_start[s]`` per source rank; ``counts[s, d]``; per destination d
    ``recv_src[d]`` / ``recv_idx[d]`` (source rank and source row of every received row, in
    receive order) and ``perm[d]`` (stable cell sort of what arrived: slot -> receive row)"""
    P = len(pos_per_rank)
    send_idx, send_start = [], []
    for s in range(P):
        a, b = route(pos_per_rank[s], nmesh, boxsize, P)
        send_idx.append(a)
        send_start.append(b)
    counts = numpy.array([[send_start[s][d + 1] - send_start[s][d] for d in range(P)] for s in range(P)], dtype='i8')
    recv_src, recv_idx, perm = [], [], []
    for d in range(P):
        src = numpy.concatenate([numpy.full(counts[s, d], s, dtype='i8') for s in range(P)])
        idx = numpy.concatenate([send_idx[s][send_start[s][d]:send_start[s][d + 1]] for s in range(P)])
        arrived = numpy.concatenate([numpy.asarray(pos_per_rank[s], dtype='f8').reshape(-1, len(nmesh))[
            send_idx[s][send_start[s][d]:send_start[s][d + 1]]] for s in range(P)])
        keys = ghost_keys(arrived, nmesh, boxsize, P, d)
        recv_src.append(src)
        recv_idx.append(idx)
        perm.append(numpy.argsort(keys, kind='stable'))
    return dict(send_idx=send_idx, send_start=send_start, counts=counts, recv_src=recv_src,
                recv_idx=recv_idx, perm=perm)
