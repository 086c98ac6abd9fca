// gridgen.cuh -- device-side generation of the cell vertices of structured grids.
//
// The reference materialises cell polygons lazily on the host (Trees.getcell: CellBasedGrid
// /root/reference/src/trees/grids.jl:71-84; HEALPix ext/ConservativeRegriddingHealpixExt.jl:76-90,
// 159-167; RingGrids ext/ConservativeRegriddingRingGridsExt.jl:22-50; Oceananigans lon-lat
// ext/ConservativeRegriddingOceananigansExt.jl:23-60,242-264).  Uploading the flattened vertex soup
// costs more than the whole build (cfg5: 435 MB over PCIe = 7.2 ms vs a 5 ms build), so grids that
// are described by a handful of numbers are generated here, straight into the build arena, with the
// same cell conventions and field-linear order and the same operation order
// as the host generators in grids.py (the two agree to the last bit except for libm sin/cos/tan).
#pragma once
#include "common.cuh"

namespace crg {

// sin and cos of an angle in degrees, exact at multiples of 90 (Julia's sincosd, which
// UnitSphereFromGeographic relies on: poles are exactly (0,0,+-1), lon 360 == lon 0).
__device__ __forceinline__ void sincosd_dev(double x, double *s, double *c) {
    double r = fmod(x, 360.0);
    if (r < 0.0) r += 360.0;
    const double q = floor((r + 45.0) / 90.0);
    const double a = (r - 90.0 * q) * (M_PI / 180.0);
    double sa, ca;
    sincos(a, &sa, &ca);
    switch (((long long)q) & 3) {
        case 0: *s = sa; *c = ca; break;
        case 1: *s = ca; *c = -sa; break;
        case 2: *s = -sa; *c = -ca; break;
        default: *s = -ca; *c = sa; break;
    }
    *s += 0.0; *c += 0.0;   // -0.0 -> +0.0
}

__device__ __forceinline__ void geo_to_xyz(double slon, double clon, double slat, double clat, double *o) {
    o[0] = clat * clon; o[1] = clat * slon; o[2] = slat + 0.0 * clon;
}

// ---- regular lon-lat grid: cell (i, j), i (longitude) fastest, ring SW, SE, NE, NW -----------------
__global__ void __launch_bounds__(256) gen_lonlat_kernel(int64_t nlon, int64_t nlat, double lon0, double lon1, double lat0,
                                                         double lat1, int64_t c0, int64_t c1, double *__restrict__ verts) {
    const int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const int64_t i = c % nlon, j = c / nlon;
    const double lw = lon0 + (lon1 - lon0) * ((double)i / (double)nlon);
    const double le = lon0 + (lon1 - lon0) * ((double)(i + 1) / (double)nlon);
    const double ls = lat0 + (lat1 - lat0) * ((double)j / (double)nlat);
    const double ln = lat0 + (lat1 - lat0) * ((double)(j + 1) / (double)nlat);
    double sw, cw, se, ce, ss, cs, sn, cn;
    sincosd_dev(lw, &sw, &cw); sincosd_dev(le, &se, &ce);
    sincosd_dev(ls, &ss, &cs); sincosd_dev(ln, &sn, &cn);
    double *o = verts + (c - c0) * 12;
    geo_to_xyz(sw, cw, ss, cs, o);
    geo_to_xyz(se, ce, ss, cs, o + 3);
    geo_to_xyz(se, ce, sn, cn, o + 6);
    geo_to_xyz(sw, cw, sn, cn, o + 9);
}

// ---- RingGrids full grid: ring-major north -> south, pole-pinned latitude edges ---------------------------
__global__ void __launch_bounds__(256) gen_full_ring_kernel(int64_t nlon, int64_t nlat, double lon_first,
                                                            const double *__restrict__ latd, int64_t c0, int64_t c1,
                                                            double *__restrict__ verts) {
    const int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const int64_t i = c % nlon, r = c / nlon;
    const double top = r == 0 ? 90.0 : 0.5 * (latd[r - 1] + latd[r]);
    const double bot = r == nlat - 1 ? -90.0 : 0.5 * (latd[r] + latd[r + 1]);
    const double dlon = 360.0 / (double)nlon;
    const double lw = lon_first - dlon / 2 + (double)i * dlon;
    const double le = lon_first - dlon / 2 + (double)(i + 1) * dlon;
    double sw, cw, se, ce, ss, cs, sn, cn;
    sincosd_dev(lw, &sw, &cw); sincosd_dev(le, &se, &ce);
    sincosd_dev(bot, &ss, &cs); sincosd_dev(top, &sn, &cn);
    double *o = verts + (c - c0) * 12;
    geo_to_xyz(sw, cw, ss, cs, o);
    geo_to_xyz(se, ce, ss, cs, o + 3);
    geo_to_xyz(se, ce, sn, cn, o + 6);
    geo_to_xyz(sw, cw, sn, cn, o + 9);
}

// ---- RingGrids reduced grid (octahedral Gaussian O<n>): ring of rank j from either pole has a + b j points --------
// The reference has no cells for reduced grids (ext/ConservativeRegriddingRingGridsExt.jl:18-20 errors; upstream
// issue #89): these generalise the full-grid rule -- latitude band between pole-pinned mid-latitudes x longitude
// interval centred on the point -- exactly like grids.py::octahedral_gaussian_grid (BASELINE config 4).
__device__ __forceinline__ int64_t reduced_ring_offset(int64_t j, int64_t a, int64_t b) {   // cells in the first j rings
    return a * j + b * (j * (j + 1) / 2);
}
__global__ void __launch_bounds__(256) gen_reduced_ring_kernel(int64_t nlat, int64_t a, int64_t b, double lon_first,
                                                               const double *__restrict__ latd, int64_t c0, int64_t c1,
                                                               double *__restrict__ verts) {
    const int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nh = nlat / 2, half = reduced_ring_offset(nh, a, b);
    if (c >= c1) return;
    const bool south = c >= half;
    const int64_t cc = south ? 2 * half - 1 - c : c;                 // mirror: the same ring rank counted from the south pole
    // largest rn with offset(rn) <= cc: root of (b/2) r^2 + (a + b/2) r - cc = 0, then fixed up
    const double B = (double)a + 0.5 * (double)b;
    int64_t rn = b > 0 ? (int64_t)((-B + sqrt(B * B + 2.0 * (double)b * (double)cc)) / (double)b) : cc / a;
    if (rn < 0) rn = 0;
    if (rn > nh - 1) rn = nh - 1;
    while (rn > 0 && reduced_ring_offset(rn, a, b) > cc) --rn;
    while (rn < nh - 1 && reduced_ring_offset(rn + 1, a, b) <= cc) ++rn;
    const int64_t n = a + b * (rn + 1);                               // points in this ring (rank j = rn + 1)
    const int64_t r = south ? nlat - 1 - rn : rn;                     // ring index north -> south
    const int64_t start = south ? 2 * half - reduced_ring_offset(rn + 1, a, b) : reduced_ring_offset(rn, a, b);
    const int64_t i = c - start;
    const double top = r == 0 ? 90.0 : 0.5 * (latd[r - 1] + latd[r]);
    const double bot = r == nlat - 1 ? -90.0 : 0.5 * (latd[r] + latd[r + 1]);
    const double dlon = 360.0 / (double)n;
    const double lw = lon_first - dlon / 2 + (double)i * dlon;
    const double le = lw + dlon;
    double sw, cw, se, ce, ss, cs, sn, cn;
    sincosd_dev(lw, &sw, &cw); sincosd_dev(le, &se, &ce);
    sincosd_dev(bot, &ss, &cs); sincosd_dev(top, &sn, &cn);
    double *o = verts + (c - c0) * 12;
    geo_to_xyz(sw, cw, ss, cs, o);
    geo_to_xyz(se, ce, ss, cs, o + 3);
    geo_to_xyz(se, ce, sn, cn, o + 6);
    geo_to_xyz(sw, cw, sn, cn, o + 9);
}

// ---- HEALPix ------------------------------------------------------------------------------------------------
__device__ __constant__ int c_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__device__ __constant__ int c_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

__device__ __forceinline__ uint32_t compress_bits(uint64_t v) {
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)v;
}
__device__ __forceinline__ int64_t isqrt64(int64_t v) {
    int64_t r = (int64_t)sqrt((double)v + 0.5);
    while (r * r > v) --r;
    while ((r + 1) * (r + 1) <= v) ++r;
    return r;
}
// ring-order pixel -> (ix, iy, face)   (standard HEALPix ring2xyf)
__device__ __forceinline__ void ring2xyf(int64_t nside, int64_t pix, int *ix, int *iy, int *face) {
    const int64_t ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside, nl2 = 2 * nside;
    int64_t iring, iphi, kshift, nr;
    int f;
    if (pix < ncap) {
        iring = (1 + isqrt64(1 + 2 * pix)) >> 1;
        iphi = (pix + 1) - 2 * iring * (iring - 1);
        kshift = 0; nr = iring;
        f = (int)((iphi - 1) / nr);
    } else if (pix < npix - ncap) {
        const int64_t ip = pix - ncap;
        const int64_t tmp = ip / (4 * nside);
        iring = tmp + nside;
        iphi = ip - tmp * 4 * nside + 1;
        kshift = (iring + nside) & 1;
        nr = nside;
        const int64_t ire = iring - nside + 1, irm = nl2 + 2 - ire;
        const int64_t ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
        f = (ifp == ifm) ? (int)(ifp | 4) : ((ifp < ifm) ? (int)ifp : (int)(ifm + 8));
    } else {
        const int64_t ip = npix - pix;
        iring = (1 + isqrt64(2 * ip - 1)) >> 1;
        iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        kshift = 0; nr = iring;
        iring = 2 * nl2 - iring;
        f = 8 + (int)((iphi - 1) / nr);
    }
    const int64_t irt = iring - c_jrll[f] * nside + 1;
    int64_t ipt = 2 * iphi - c_jpll[f] * nr - kshift - 1;
    if (ipt >= nl2) ipt -= 8 * nside;
    *ix = (int)((ipt - irt) >> 1);
    *iy = (int)((-ipt - irt) >> 1);
    *face = f;
}
// point (x, y) in [0,1]^2 of a base face -> unit vector (HEALPix xyf2loc; same arithmetic as grids.py)
__device__ __forceinline__ void healpix_loc(double x, double y, int face, double *o) {
    const double jr = (double)c_jrll[face] - x - y;
    double nr, z, sth;
    if (jr < 1.0) {
        nr = jr;
        const double tmp = nr * nr / 3.0;
        z = 1.0 - tmp;
        sth = sqrt(fmax(tmp * (2.0 - tmp), 0.0));
    } else if (jr > 3.0) {
        nr = 4.0 - jr;
        const double tmp = nr * nr / 3.0;
        z = tmp - 1.0;
        sth = sqrt(fmax(tmp * (2.0 - tmp), 0.0));
    } else {
        nr = 1.0;
        z = (2.0 - jr) * 2.0 / 3.0;
        sth = sqrt(fmax((1.0 - z) * (1.0 + z), 0.0));
    }
    double t = (double)c_jpll[face] * nr + x - y;
    if (t < 0.0) t += 8.0;
    if (t >= 8.0) t -= 8.0;
    const double phi = nr < 1e-15 ? 0.0 : (0.25 * M_PI * t) / nr;
    double sp, cp;
    sincos(phi, &sp, &cp);
    o[0] = sth * cp; o[1] = sth * sp; o[2] = z;
}
// corners N, W, S, E (CCW from outside) = Healpix.boundariesRing(res, pix, 1)
__global__ void __launch_bounds__(256) gen_healpix_kernel(int64_t nside, int nested, int64_t c0, int64_t c1,
                                                          double *__restrict__ verts) {
    const int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    int ix, iy, face;
    if (nested) {
        const int64_t npface = nside * nside;
        face = (int)(c / npface);
        const uint64_t p = (uint64_t)(c % npface);
        ix = (int)compress_bits(p);
        iy = (int)compress_bits(p >> 1);
    } else {
        ring2xyf(nside, c, &ix, &iy, &face);
    }
    const double x0 = (double)ix / (double)nside, x1 = (double)(ix + 1) / (double)nside;
    const double y0 = (double)iy / (double)nside, y1 = (double)(iy + 1) / (double)nside;
    double *o = verts + (c - c0) * 12;
    healpix_loc(x1, y1, face, o);       // N
    healpix_loc(x0, y1, face, o + 3);   // W
    healpix_loc(x0, y0, face, o + 6);   // S
    healpix_loc(x1, y0, face, o + 9);   // E
}

// ---- equiangular gnomonic cubed sphere C<n>: 6 panels, panel-major, i fastest ----------------------------
__device__ __forceinline__ double cs_coord(int64_t i, int64_t n) {
    if (i == 0) return -1.0;
    if (i == n) return 1.0;
    if (2 * i == n) return 0.0;
    return tan(-M_PI / 4 + (M_PI / 2) * ((double)i / (double)n));
}
__global__ void __launch_bounds__(256) gen_cubed_sphere_kernel(int64_t n, int64_t c0, int64_t c1, double *__restrict__ verts) {
    const int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const int panel = (int)(c / (n * n));
    const int64_t k = c % (n * n), i = k % n, j = k / n;
    const int64_t ci[4] = {i, i + 1, i + 1, i}, cj[4] = {j, j, j + 1, j + 1};
    double *o = verts + (c - c0) * 12;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double X = cs_coord(ci[q], n), Y = cs_coord(cj[q], n);
        double px, py, pz;
        switch (panel) {
            case 0: px = 1.0; py = X; pz = Y; break;
            case 1: px = -X; py = 1.0; pz = Y; break;
            case 2: px = -1.0; py = -X; pz = Y; break;
            case 3: px = X; py = -1.0; pz = Y; break;
            case 4: px = -Y; py = X; pz = 1.0; break;
            default: px = Y; py = X; pz = -1.0; break;
        }
        const double nrm = sqrt(px * px + py * py + pz * pz);
        o[3 * q] = px / nrm; o[3 * q + 1] = py / nrm; o[3 * q + 2] = pz / nrm;
    }
}

}  // namespace crg
