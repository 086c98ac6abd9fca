// Host build of the device arithmetic in csrc/clipfast.cuh (same source, CRG_HD), driven pair by pair from
// tests/test_clipfast_host.py against the oracle.  Test infrastructure only.
#include "../../conservativeregridding.jl_b200/csrc/clipfast.cuh"

extern "C" void cf_pairs_host(const double *sverts, const unsigned char *sflip, const double *cverts,
                              const unsigned char *cflip, const long long *si, const long long *ci, long long n,
                              double *area, int *kind, int stored_vertex_as_P) {
    for (long long k = 0; k < n; ++k) {
        double s[4][3], c[4][3], nrm[4][3], cor[4][3];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) { s[i][j] = sverts[si[k] * 12 + i * 3 + j]; c[i][j] = cverts[ci[k] * 12 + i * 3 + j]; }
        crg::cf_quad_normals(c, cflip && cflip[ci[k]], nrm);
        crg::cf_quad_corners(c, nrm, cor);
        area[k] = crg::cf_pair_area(s, sflip && sflip[si[k]], nrm, stored_vertex_as_P ? cverts + ci[k] * 12 : &cor[0][0], &kind[k]);
    }
}
