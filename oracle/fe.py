"""Reference-element tables for the Q2-Q1 Taylor-Hood pair (oracle; test infrastructure).

Follows the constants restated in SURVEY.md App. B and the deal.II conventions
used at /root/reference/source/simulator/solver/matrix_free_operators.cc:384-385
(FEEvaluation<dim,degree_v,degree_v+1,dim> and <dim,degree_v-1,degree_v+1,1>:
velocity and pressure share the QGauss<1>(degree_v+1) = 3-point rule) and
:795 (mass operator: FEEvaluation<dim,degree_p,degree_p+2,1>, again 3 points).
"""
import itertools
import numpy as np

# 3-point Gauss rule on [0,1]
GAUSS_X = np.array([0.5 - 0.5 * np.sqrt(3.0 / 5.0), 0.5, 0.5 + 0.5 * np.sqrt(3.0 / 5.0)])
GAUSS_W = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])


def q2_shape(x):
    """Q2 Lagrange basis on nodes {0, 1/2, 1}: values and derivatives, shape (len(x), 3)."""
    x = np.asarray(x, dtype=float)
    v = np.stack([(1 - x) * (1 - 2 * x), 4 * x * (1 - x), x * (2 * x - 1)], axis=-1)
    d = np.stack([4 * x - 3, 4 - 8 * x, 4 * x - 1], axis=-1)
    return v, d


def q1_shape(x):
    x = np.asarray(x, dtype=float)
    v = np.stack([1 - x, x], axis=-1)
    d = np.stack([-np.ones_like(x), np.ones_like(x)], axis=-1)
    return v, d


def gauss(n):
    """n-point Gauss rule on [0,1]."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


N2, D2 = q2_shape(GAUSS_X)   # [q][i]
N1, D1 = q1_shape(GAUSS_X)

# 1D prolongation parent -> child-patch nodes
P2_1D = np.array([[1, 0, 0], [3 / 8, 3 / 4, -1 / 8], [0, 1, 0], [-1 / 8, 3 / 4, 3 / 8], [0, 0, 1]], dtype=float)
P1_1D = np.array([[1, 0], [0.5, 0.5], [0, 1]], dtype=float)


def lex_nodes(npd, dim):
    """Lexicographic (x fastest) list of node multi-indices for npd nodes per direction."""
    return [tuple(reversed(t)) for t in itertools.product(range(npd), repeat=dim)]


def fe_q_local_order(degree, dim):
    """deal.II FE_Q cell-local ordering (vertices, lines, quads, hex interior) expressed as a
    list of lexicographic node multi-indices.

    vertex v = x + 2y + 4z; 3D lines: 0-3 bottom face {x=0 along y, x=1 along y, y=0 along x,
    y=1 along x}, 4-7 same on top, 8-11 vertical; quads (faces) x0,x1,y0,y1,z0,z1; 2D lines
    x=0, x=1, y=0, y=1.
    """
    npd = degree + 1
    hi = npd - 1
    order = []
    # vertices
    for v in range(2 ** dim):
        order.append(tuple(hi * ((v >> d) & 1) for d in range(dim)))
    if degree == 1:
        return order
    assert degree == 2
    mid = 1
    if dim == 2:
        order += [(0, mid), (hi, mid), (mid, 0), (mid, hi)]       # lines
        order += [(mid, mid)]                                     # quad interior
    else:
        for z in (0, hi):
            order += [(0, mid, z), (hi, mid, z), (mid, 0, z), (mid, hi, z)]
        order += [(0, 0, mid), (hi, 0, mid), (0, hi, mid), (hi, hi, mid)]
        order += [(0, mid, mid), (hi, mid, mid), (mid, 0, mid), (mid, hi, mid), (mid, mid, 0), (mid, mid, hi)]
        order += [(mid, mid, mid)]
    assert len(order) == npd ** dim and len(set(order)) == len(order)
    return order


def tensor_tables(dim):
    """Shape values/gradients at the 3^dim Gauss points for Q2 and Q1 scalars on the unit cell.

    returns dict with
      w[q]                 quadrature weights (reference cell)
      xq[q, d]             quadrature points
      phi2[q, i], dphi2[q, i, d]   (i lexicographic over 3^dim nodes)
      phi1[q, i], dphi1[q, i, d]   (i lexicographic over 2^dim nodes)
    q is lexicographic over the Gauss points (x fastest).
    """
    qs = lex_nodes(3, dim)
    n2 = lex_nodes(3, dim)
    n1 = lex_nodes(2, dim)
    w = np.array([np.prod([GAUSS_W[q[d]] for d in range(dim)]) for q in qs])
    xq = np.array([[GAUSS_X[q[d]] for d in range(dim)] for q in qs])

    def tab(N, D, nodes):
        phi = np.zeros((len(qs), len(nodes)))
        dphi = np.zeros((len(qs), len(nodes), dim))
        for a, q in enumerate(qs):
            for b, i in enumerate(nodes):
                phi[a, b] = np.prod([N[q[d], i[d]] for d in range(dim)])
                for e in range(dim):
                    dphi[a, b, e] = np.prod([(D if d == e else N)[q[d], i[d]] for d in range(dim)])
        return phi, dphi

    phi2, dphi2 = tab(N2, D2, n2)
    phi1, dphi1 = tab(N1, D1, n1)
    return dict(w=w, xq=xq, phi2=phi2, dphi2=dphi2, phi1=phi1, dphi1=dphi1)
