// oracle/ref_permuto_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" access to the REAL reference lattice: this file is compiled together with
// /root/reference/CRF/src/permutohedral.cpp (unmodified, in place) against the Eigen
// stand-in in oracle/eigen_shim, producing oracle/_ref/libpermuto_ref.so.  The class
// members are `protected` (CRF/include/permutohedral.h:43-57), hence the subclass.
#include "permutohedral.h"

class RefLattice : public Permutohedral {
public:
    int M() const { return M_; }
    int N() const { return N_; }
    int d() const { return d_; }
    const int *offset() const { return offset_.data(); }
    const int *rank() const { return rank_.data(); }
    const float *bary() const { return barycentric_.data(); }
    int n1(int j, int i) const { return blur_neighbors_[(size_t)j * M_ + i].n1; }
    int n2(int j, int i) const { return blur_neighbors_[(size_t)j * M_ + i].n2; }
    void seq(float *out, const float *in, int vs) const { seqCompute(out, in, vs, false); }
    void sse(float *out, const float *in, int vs) const { sseCompute(out, in, vs, false); }
};

extern "C" {
// feature: d x N column-major (feature[k*d+j]) exactly like Eigen's MatrixXf(d, N).
void *ref_lattice_init(const float *feature, int d, int N) {
    MatrixXf f(d, N);
    for (size_t i = 0; i < (size_t)d * N; i++) f.data()[i] = feature[i];
    RefLattice *L = new RefLattice();
    L->init(f);
    return L;
}
void ref_lattice_free(void *p) { delete (RefLattice *)p; }
int ref_lattice_M(void *p) { return ((RefLattice *)p)->M(); }
const int *ref_lattice_offset(void *p) { return ((RefLattice *)p)->offset(); }
const int *ref_lattice_rank(void *p) { return ((RefLattice *)p)->rank(); }
const float *ref_lattice_bary(void *p) { return ((RefLattice *)p)->bary(); }
void ref_lattice_neighbors(void *p, int *n1, int *n2) {
    RefLattice *L = (RefLattice *)p;
    for (int j = 0; j <= L->d(); j++)
        for (int i = 0; i < L->M(); i++) {
            n1[(size_t)j * L->M() + i] = L->n1(j, i);
            n2[(size_t)j * L->M() + i] = L->n2(j, i);
        }
}
void ref_lattice_seq_compute(void *p, float *out, const float *in, int vs) {
    ((RefLattice *)p)->seq(out, in, vs);
}
void ref_lattice_sse_compute(void *p, float *out, const float *in, int vs) {
    ((RefLattice *)p)->sse(out, in, vs);
}
// the public dispatching entry point, CRF/src/permutohedral.cpp:596-604
void ref_lattice_compute(void *p, float *out, const float *in, int vs, int N) {
    MatrixXf mi(vs, N), mo(vs, N);
    for (size_t i = 0; i < (size_t)vs * N; i++) mi.data()[i] = in[i];
    ((RefLattice *)p)->compute(mo, mi, false);
    for (size_t i = 0; i < (size_t)vs * N; i++) out[i] = mo.data()[i];
}
}
