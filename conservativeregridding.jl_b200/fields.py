"""Analytic test fields (Valcke et al. 2022), restating the formulas of
/root/reference/src/utils/example_data.jl:22-24,39-41,52-54,75-114,132-154.
Used for synthetic bench inputs and conservation tests; all take (lon, lat) in degrees.
"""
from __future__ import annotations

import numpy as np

from .grids import sincosd


def longitude_field(lon, lat):
    """``LongitudeField`` (example_data.jl:22-24)."""
    return np.asarray(lon, dtype=np.float64) + 0.0 * np.asarray(lat)


def sinusoid_field(lon, lat, dp_length=1.2 * np.pi, coef=2.0, coefmult=1.0):
    """``SinusoidField`` (example_data.jl:39-41)."""
    lon = np.radians(lon)
    lat = np.radians(lat)
    return coefmult * (coef - np.cos(np.pi * np.arccos(np.clip(np.cos(lon) * np.cos(lat), -1, 1)) / dp_length))


def harmonic_field(lon, lat):
    """``HarmonicField`` (example_data.jl:52-54)."""
    return 2.0 + (np.sin(2 * np.radians(lat)) ** 16) * np.cos(16 * np.radians(lon))


def gulf_stream_field(lon, lat, dp_length=1.2 * np.pi, coef=2.0, gf_coef=1.0, gf_ori_lon=-80.0,
                      gf_ori_lat=25.0, gf_end_lon=-1.8, gf_end_lat=50.0, gf_dmp_lon=-25.5,
                      gf_dmp_lat=55.5):
    """``GulfStreamField`` (example_data.jl:75-114)."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    k = np.pi / 180.0
    dr0 = np.hypot((gf_end_lon - gf_ori_lon) * k, (gf_end_lat - gf_ori_lat) * k)
    dr1 = np.hypot((gf_dmp_lon - gf_ori_lon) * k, (gf_dmp_lat - gf_ori_lat) * k)
    res = coef - np.cos(np.pi * np.arccos(np.clip(np.cos(lat * k) * np.cos(lon * k), -1, 1)) / dp_length)
    per = np.where(lon > 180.0, lon - 360.0, np.where(lon < -180.0, lon + 360.0, lon))
    dx = (per - gf_ori_lon) * k
    dy = (lat - gf_ori_lat) * k
    dr = np.hypot(dx, dy)
    dth = np.arctan2(dy, dx)
    dc = np.full_like(dr, 1.3 * gf_coef)
    mid = (dr > dr1) & (dr <= dr0)
    dc = np.where(mid, dc * np.cos(np.pi * 0.5 * (dr - dr1) / (dr0 - dr1)), dc)
    dc = np.where(dr > dr0, 0.0, dc)
    res = res + (np.maximum(1000.0 * np.sin(0.4 * (0.5 * dr + dth) + 0.007 * np.cos(50.0 * dth)
                                             + 0.37 * np.pi), 999.0) - 999.0) * dc
    return res


def vortex_field(lon, lat, lon0_rad=5.5, lat0_rad=0.2, r0=3.0, d=5.0, t=6.0):
    """``VortexField`` (example_data.jl:132-154)."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    sin_c, cos_c = np.sin(lat0_rad), np.cos(lat0_rad)
    sin_lat, cos_lat = sincosd(lat)
    trm = cos_lat * np.cos(np.radians(lon) - lon0_rad)
    X = sin_c * trm - cos_c * sin_lat
    Y = cos_lat * np.sin(np.radians(lon) - lon0_rad)
    Z = sin_c * sin_lat + cos_c * trm
    dlon = np.arctan2(Y, X)
    dlon = np.where(dlon < 0, dlon + 2 * np.pi, dlon)
    dlat = np.arcsin(np.clip(Z, -1, 1))
    rho = r0 * np.cos(dlat)
    vt = 3 * np.sqrt(3) / 2 / np.cosh(rho) ** 2 * np.tanh(rho)
    with np.errstate(divide="ignore", invalid="ignore"):
        omega = np.where(rho == 0, 0.0, vt / np.where(rho == 0, 1.0, rho))
    return 2 * (1 + np.tanh(rho / d * np.sin(dlon - omega * t)))


EXAMPLE_FIELDS = {
    "longitude": longitude_field,
    "sinusoid": sinusoid_field,
    "harmonic": harmonic_field,
    "gulf_stream": gulf_stream_field,
    "vortex": vortex_field,
}
