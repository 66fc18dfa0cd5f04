// Device-side flattened form of bgp_kernel_spec: one entry per (term, active dim) "feature".
#pragma once
#include "common.cuh"

namespace bgp {

struct DevSpec {
    int nfeat;
    int ftype[16];      // term type of feature f
    int fdim[16];       // column of X
    int flast[16];      // 1 when f is the last feature of its term
    double fscale[16];  // multiply... see prescale(): RBF/MATERN 1/l (applied as division), PERIODIC pi/p
    double faux[16];    // PERIODIC: 1/l
    double fos[16];     // outputscale of the term (valid at flast)
    double noise;
};

static inline int make_devspec(const bgp_kernel_spec* s, DevSpec* d) {
    if (!s || s->nterms < 1 || s->nterms > BGP_MAX_TERMS) return BGP_E_SPEC;
    int f = 0;
    for (int t = 0; t < s->nterms; t++) {
        const bgp_term& T = s->terms[t];
        if (T.type < BGP_WIENER || T.type > BGP_PERIODIC) return BGP_E_SPEC;
        const int nd = (T.type == BGP_WIENER) ? 1 : T.ndims;
        if (nd < 1 || nd > BGP_MAX_DIMS || f + nd > 16) return BGP_E_SPEC;
        for (int k = 0; k < nd; k++, f++) {
            if (T.dims[k] < 0) return BGP_E_SPEC;
            d->ftype[f] = T.type;
            d->fdim[f] = T.dims[k];
            d->flast[f] = (k == nd - 1);
            d->fos[f] = T.outputscale;
            d->faux[f] = 0.0;
            if (T.type == BGP_WIENER) d->fscale[f] = 1.0;
            else if (T.type == BGP_PERIODIC) {
                if (!(T.period[k] > 0.0) || !(T.lengthscale[k] > 0.0)) return BGP_E_SPEC;
                d->fscale[f] = 3.14159265358979323846 / T.period[k];
                d->faux[f] = 1.0 / T.lengthscale[k];
            } else {
                if (!(T.lengthscale[k] > 0.0)) return BGP_E_SPEC;
                d->fscale[f] = T.lengthscale[k];
            }
        }
    }
    d->nfeat = f;
    d->noise = s->noise;
    return 0;
}

__device__ __forceinline__ double prescale(const DevSpec& sp, int f, double x) {
    const int ty = sp.ftype[f];
    if (ty == BGP_WIENER) return x;
    if (ty == BGP_PERIODIC) return x * sp.fscale[f];
    return x / sp.fscale[f];
}


}  // namespace bgp
