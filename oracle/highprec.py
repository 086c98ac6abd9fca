"""50-digit (mpmath) intersection areas of spherical polygon pairs -- an arithmetic-independent
check of the oracle's Float64 clip + area (TEST INFRASTRUCTURE; see crg_oracle.c header).

Same published algorithm (Sutherland-Hodgman against great-circle half-spaces), but the area is
taken with a DIFFERENT formula -- Girard's theorem, sum of interior angles - (n-2) pi, which is
what the reference says GeometryOps uses (ext/ConservativeRegriddingClimaCoreExt.jl:345) -- so
agreement validates both the clip and the excess formula of the oracle.
"""
import mpmath as mp

mp.mp.dps = 50


def _v(p):
    return [mp.mpf(float(x)) for x in p]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _norm(a):
    n = mp.sqrt(_dot(a, a))
    return [x / n for x in a]


def girard_area(poly):
    """Signed spherical excess of a CCW polygon from its interior angles."""
    pts = []
    for p in poly:                      # drop exact duplicates (degenerate SH output)
        if not pts or any(abs(p[k] - pts[-1][k]) > mp.mpf(10) ** -45 for k in range(3)):
            pts.append(p)
    if len(pts) > 1 and all(abs(pts[0][k] - pts[-1][k]) <= mp.mpf(10) ** -45 for k in range(3)):
        pts.pop()
    n = len(pts)
    if n < 3:
        return mp.mpf(0)
    total = mp.mpf(0)
    for i in range(n):
        a, b, c = pts[i - 1], pts[i], pts[(i + 1) % n]
        # interior angle at b between the tangents towards a and c
        ta = _cross(_cross(b, a), b)
        tc = _cross(_cross(b, c), b)
        ang = mp.atan2(_dot(b, _cross(tc, ta)), _dot(ta, tc))
        if ang < 0:
            ang += 2 * mp.pi
        total += ang
    return total - (n - 2) * mp.pi


def clip(subj, clipper):
    cur = [_v(p) for p in subj]
    cl = [_v(p) for p in clipper]
    for k in range(len(cl)):
        u, v = cl[k], cl[(k + 1) % len(cl)]
        n = _cross(u, v)
        if _dot(n, n) == 0:
            continue
        out = []
        m = len(cur)
        for i in range(m):
            p, q = cur[i], cur[(i + 1) % m]
            dp, dq = _dot(n, p), _dot(n, q)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append(_norm([p[j] + t * (q[j] - p[j]) for j in range(3)]))
            if dq >= 0:
                out.append(q)
        cur = out
        if not cur:
            break
    return cur


def intersection_area(p1, p2):
    """High-precision area of the intersection of two CCW convex spherical polygons (float)."""
    poly = clip(p1, p2)
    if len(poly) < 3:
        return 0.0
    return float(girard_area(poly))


def polygon_area(p):
    return float(girard_area([_v(q) for q in p]))


# ---------------------------------------------------------------------------------------------
# An algorithm-independent check: NO Sutherland-Hodgman.  The intersection of two convex spherical
# polygons is the convex hull of (vertices of A inside B) + (vertices of B inside A) + (proper crossings
# of an edge of A with an edge of B); the points are ordered by angle around their centroid and the area
# is Girard's excess.  With 50 digits the result is, for all practical purposes, the exact area for the
# exact Float64 input vertices -- so it also pins the RESULT of the Float64 clip (oracle and CUDA), not only
# its arithmetic.  Used by tests/golden/make_highprec_pairs.py.
# ---------------------------------------------------------------------------------------------

def _edge_normals(P):
    return [_cross(P[i], P[(i + 1) % len(P)]) for i in range(len(P))]


def _orient_ccw(P):
    """Return the ring counter-clockwise (seen from outside), dropping exact duplicate vertices."""
    Q = []
    for p in P:
        if not Q or any(p[k] != Q[-1][k] for k in range(3)):
            Q.append(p)
    if len(Q) > 1 and all(Q[0][k] == Q[-1][k] for k in range(3)):
        Q.pop()
    if len(Q) < 3:
        return Q
    if girard_signed(Q) < 0:
        Q = Q[::-1]
    return Q


def girard_signed(pts):
    """Signed area by the fan of triangle excesses (only used to find the orientation)."""
    a = pts[0]
    tot = mp.mpf(0)
    for i in range(1, len(pts) - 1):
        b, c = pts[i], pts[i + 1]
        tot += 2 * mp.atan2(_dot(a, _cross(b, c)), 1 + _dot(a, b) + _dot(b, c) + _dot(c, a))
    return tot


def _exact_ints(points):
    """Float64 coordinates as integers over one common power-of-two denominator (exact)."""
    ratios = [[float(x).as_integer_ratio() for x in p] for p in points]
    K = max(d.bit_length() - 1 for p in ratios for (_, d) in p)
    return [[n << (K - (d.bit_length() - 1)) for (n, d) in p] for p in ratios]


def _icross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _idot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def independent_intersection_area(p1, p2, return_points=False):
    """Area of the intersection of two convex spherical polygons WITHOUT Sutherland-Hodgman (see above).
    Every in/out and straddling decision is taken in EXACT integer arithmetic on the Float64 inputs (so
    coincident edges and shared vertices -- nested or identical grids -- are classified exactly, `>= 0` =
    inside like the reference's clip); coordinates of crossings and the area are 50-digit."""
    p1 = [list(map(float, p)) for p in p1]
    p2 = [list(map(float, p)) for p in p2]

    def dedupe(P):
        Q = []
        for p in P:
            if not Q or p != Q[-1]:
                Q.append(p)
        if len(Q) > 1 and Q[0] == Q[-1]:
            Q.pop()
        return Q
    p1, p2 = dedupe(p1), dedupe(p2)
    if len(p1) < 3 or len(p2) < 3:
        return (0.0, []) if return_points else 0.0
    # orientation from the 50-digit signed area; rings are made counter-clockwise
    if girard_signed([_v(p) for p in p1]) < 0:
        p1 = p1[::-1]
    if girard_signed([_v(p) for p in p2]) < 0:
        p2 = p2[::-1]
    ints = _exact_ints(p1 + p2)
    IA, IB = ints[:len(p1)], ints[len(p1):]
    # (directions only: Float64 unit vectors are off the sphere by ~1e-16, computed crossings are not)
    A = [_norm(_v(p)) for p in p1]
    B = [_norm(_v(p)) for p in p2]
    na, nb = len(A), len(B)
    inA = [_icross(IA[i], IA[(i + 1) % na]) for i in range(na)]
    inB = [_icross(IB[j], IB[(j + 1) % nb]) for j in range(nb)]
    # exact signed distances: DA[j][i] = nB_j . a_i  (a_i inside B <=> all >= 0);  DB[i][j] = nA_i . b_j
    DA = [[_idot(inB[j], IA[i]) for i in range(na)] for j in range(nb)]
    DB = [[_idot(inA[i], IB[j]) for j in range(nb)] for i in range(na)]
    pts = []
    for i in range(na):
        if all(DA[j][i] >= 0 for j in range(nb)):
            pts.append(A[i])
    for j in range(nb):
        if all(DB[i][j] >= 0 for i in range(na)):
            pts.append(B[j])
    nA = [_cross(A[i], A[(i + 1) % na]) for i in range(na)]
    nB = [_cross(B[j], B[(j + 1) % nb]) for j in range(nb)]
    for i in range(na):
        i2 = (i + 1) % na
        for j in range(nb):
            j2 = (j + 1) % nb
            # arc a_i a_i2 straddles the great circle of b_j b_j2 and vice versa (arcs are shorter than pi)
            if (DA[j][i] >= 0) == (DA[j][i2] >= 0) or (DB[i][j] >= 0) == (DB[i][j2] >= 0):
                continue
            d = _cross(nA[i], nB[j])
            if _dot(d, d) == 0:
                continue
            x = _norm(d)
            mid = [A[i][k] + A[i2][k] for k in range(3)]
            if _dot(x, mid) < 0:
                x = [-t for t in x]
            pts.append(x)
    # distinct points (a vertex on the other polygon's edge also shows up as a crossing)
    tol = mp.mpf(10) ** -40
    uniq = []
    for p in pts:
        if not any(all(abs(p[k] - q[k]) <= tol for k in range(3)) for q in uniq):
            uniq.append(p)
    if len(uniq) < 3:
        return (0.0, uniq) if return_points else 0.0
    c = _norm([sum(p[k] for p in uniq) for k in range(3)])
    ref = [uniq[0][k] - c[k] * _dot(uniq[0], c) for k in range(3)]
    if _dot(ref, ref) < tol:
        ref = [uniq[1][k] - c[k] * _dot(uniq[1], c) for k in range(3)]
    e1 = _norm(ref)
    e2 = _cross(c, e1)
    order = sorted(range(len(uniq)), key=lambda k: mp.atan2(_dot(uniq[k], e2), _dot(uniq[k], e1)))
    poly = [uniq[k] for k in order]                   # counter-clockwise around c seen from outside
    # points on one great circle (cells that only touch along an edge): no area, and the interior angles
    # Girard's formula needs are undefined there
    area = mp.mpf(0) if abs(girard_signed(poly)) < mp.mpf(10) ** -35 else girard_area(poly)
    if return_points:
        return float(area), poly
    return float(area)
