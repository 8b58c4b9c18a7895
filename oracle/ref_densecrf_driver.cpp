// oracle/ref_densecrf_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" access to the REAL reference CRF: compiled together with the reference's own, unmodified
// CRF/src/{densecrf,pairwise,labelcompatibility,unary,util,permutohedral,densecrf_wrapper}.cpp (the list of
// CRF/setup.py:17-25) against the Eigen stand-in in oracle/eigen_shim into oracle/_ref/libdensecrf_ref.so.
// The functions map 1:1 onto the methods of DenseCRFWrapper (CRF/include/densecrf_wrapper.h:3-28), the class
// the reference's Cython extension binds.
#include "densecrf_wrapper.h"

extern "C" {
void *ref_densecrf_create(int W, int H, int nlabels) { return new DenseCRFWrapper(W, H, nlabels); }
void ref_densecrf_destroy(void *p) { delete (DenseCRFWrapper *)p; }
void ref_densecrf_set_unary_energy(void *p, float *unary_costs) { ((DenseCRFWrapper *)p)->set_unary_energy(unary_costs); }
void ref_densecrf_add_pairwise_energy(void *p, float w1, float ta1, float ta2, float tb1, float tb2, float tb3,
                                      float w2, float tg1, float tg2, unsigned char *im) {
    ((DenseCRFWrapper *)p)->add_pairwise_energy(w1, ta1, ta2, tb1, tb2, tb3, w2, tg1, tg2, im);
}
void ref_densecrf_inference(void *p, int n_iters, float *probs_out) { ((DenseCRFWrapper *)p)->inference(n_iters, probs_out); }
void ref_densecrf_map(void *p, int n_iters, int *labels) { ((DenseCRFWrapper *)p)->map(n_iters, labels); }
}
