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
