/*
 * oracle/crf_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99) of the dense-CRF mean-field inference path of
 * speedinghzl/DSRG's vendored Kraehenbuehl-2013 code.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load it.
 *
 * Parity status: PINNED against the reference's own code.  The reference ships no golden
 * vectors for this path and real Eigen3 is an absent, un-vendored dependency, but its CRF
 * sources compile unmodified, in place, against a small Eigen stand-in (oracle/eigen_shim):
 *   - oracle/_ref/libpermuto_ref.so  = CRF/src/permutohedral.cpp: the lattice part of this
 *     file is checked bit-for-bit (offsets, barycentrics, ranks, blur neighbours, seq/sse
 *     compute) against it;
 *   - oracle/_ref/libdensecrf_ref.so = the whole source list of CRF/setup.py:17-25 behind
 *     DenseCRFWrapper: inference() and map() of this file are BIT-IDENTICAL to it on every
 *     test case (tests/test_oracle_golden.py), i.e. features, norm, Potts sign, kernel order,
 *     mean-field loop and softmax structure are the reference's.
 * What stays unpinned is Eigen's own numerics: its vectorised exp() and packet-wise sum()
 * are expf() and a sequential sum both here and in the stand-in (<= few ulp, far inside 1e-4).
 *
 * Build flags matter: -O2 -ffp-contract=off and NO -march (the reference build,
 * CRF/setup.py:14-33, passes no arch flags => SSE2 float math, no FMA, MXCSR
 * round-to-nearest-even).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Hash table of short keys: CRF/src/permutohedral.cpp:54-131          */
/* ------------------------------------------------------------------ */
typedef struct {
    size_t key_size, filled, capacity;
    short *keys; /* (capacity/2+10)*key_size */
    int *table;  /* capacity, -1 = empty */
} OHash;

static size_t ohash_hash(const OHash *h, const short *k) { /* :80-87 */
    size_t r = 0;
    for (size_t i = 0; i < h->key_size; i++) {
        r += (size_t)(long)k[i];
        r *= 1664525;
    }
    return r;
}

static void ohash_init(OHash *h, int key_size, int n_elements) { /* :89-90 */
    h->key_size = (size_t)key_size;
    h->filled = 0;
    h->capacity = 2 * (size_t)n_elements;
    h->keys = (short *)calloc((h->capacity / 2 + 10) * h->key_size, sizeof(short));
    h->table = (int *)malloc(h->capacity * sizeof(int));
    for (size_t i = 0; i < h->capacity; i++) h->table[i] = -1;
}

static void ohash_free(OHash *h) {
    free(h->keys);
    free(h->table);
}

static void ohash_grow(OHash *h) { /* :59-79 */
    size_t old_capacity = h->capacity;
    h->capacity *= 2;
    short *nk = (short *)calloc((old_capacity + 10) * h->key_size, sizeof(short));
    memcpy(nk, h->keys, (old_capacity / 2 + 10) * h->key_size * sizeof(short));
    int *nt = (int *)malloc(h->capacity * sizeof(int));
    for (size_t i = 0; i < h->capacity; i++) nt[i] = -1;
    int *old_table = h->table;
    free(h->keys);
    h->keys = nk;
    h->table = nt;
    for (size_t i = 0; i < old_capacity; i++)
        if (old_table[i] >= 0) {
            int e = old_table[i];
            size_t hh = ohash_hash(h, h->keys + (size_t)e * h->key_size) % h->capacity;
            for (; h->table[hh] >= 0; hh = hh < h->capacity - 1 ? hh + 1 : 0)
                ;
            h->table[hh] = e;
        }
    free(old_table);
}

static int ohash_find(OHash *h, const short *k, int create) { /* :98-126 */
    if (2 * h->filled >= h->capacity) ohash_grow(h);
    size_t hh = ohash_hash(h, k) % h->capacity;
    for (;;) {
        int e = h->table[hh];
        if (e == -1) {
            if (create) {
                for (size_t i = 0; i < h->key_size; i++)
                    h->keys[h->filled * h->key_size + i] = k[i];
                h->table[hh] = (int)h->filled;
                return (int)h->filled++;
            }
            return -1;
        }
        int good = 1;
        for (size_t i = 0; i < h->key_size && good; i++)
            if (h->keys[(size_t)e * h->key_size + i] != k[i]) good = 0;
        if (good) return e;
        hh++;
        if (hh == h->capacity) hh = 0;
    }
}

/* ------------------------------------------------------------------ */
/* Permutohedral lattice                                                */
/* ------------------------------------------------------------------ */
typedef struct {
    int N, M, d;
    int *offset;  /* (d+1)*(N+16), pixel-major: offset[i*(d+1)+r]  */
    float *bary;  /* same indexing                                 */
    int *rank;    /* same indexing                                 */
    int *n1, *n2; /* (d+1)*M: n1[j*M+i]                            */
} OLattice;

void oracle_lattice_free(OLattice *L) {
    if (!L) return;
    free(L->offset);
    free(L->bary);
    free(L->rank);
    free(L->n1);
    free(L->n2);
    free(L);
}

/*
 * Permutohedral::init, SSE variant with the non-SSE4.1 rounding branch
 * (CRF/src/permutohedral.cpp:140-321; the #ifndef __SSE4_1__ lines :185-188,
 * :215-217 are the ones a flag-less distutils build compiles).  The 4-wide SSE
 * lanes are restated lane by lane in scalar float arithmetic (same IEEE single
 * ops, no contraction); lanes beyond N carry feature 0 and DO insert vertices
 * (:196, :261-275).
 * feature is d x N column-major (Eigen MatrixXf): feature[k*d + j] = f_j(pixel k).
 */
OLattice *oracle_lattice_init(const float *feature, int d, int N) {
    OLattice *L = (OLattice *)calloc(1, sizeof(OLattice));
    L->N = N;
    L->d = d;
    OHash ht;
    ohash_init(&ht, d, N); /* :145 */
    const int B = 4;       /* blocksize :147 */
    const float invdplus1 = 1.0f / (d + 1); /* :148 */
    const float dplus1 = (float)(d + 1);    /* :149 */
    size_t sz = (size_t)(d + 1) * (size_t)(N + 16);
    L->offset = (int *)calloc(sz, sizeof(int));   /* :154-155 */
    L->bary = (float *)calloc(sz, sizeof(float)); /* :156-157 */
    L->rank = (int *)calloc(sz, sizeof(int));     /* :158 */

    float *scale_factor = (float *)malloc(sizeof(float) * d);
    float *f = (float *)malloc(sizeof(float) * d * B);
    float *elevated = (float *)malloc(sizeof(float) * (d + 1) * B);
    float *rem0 = (float *)malloc(sizeof(float) * (d + 1) * B);
    float *rank = (float *)malloc(sizeof(float) * (d + 1) * B);
    float *barycentric = (float *)malloc(sizeof(float) * (d + 2) * B);
    short *canonical = (short *)malloc(sizeof(short) * (d + 1) * (d + 1));
    short *key = (short *)malloc(sizeof(short) * (d + 1));

    for (int i = 0; i <= d; i++) { /* :171-176 */
        for (int j = 0; j <= d - i; j++) canonical[i * (d + 1) + j] = (short)i;
        for (int j = d - i + 1; j <= d; j++) canonical[i * (d + 1) + j] = (short)(i - (d + 1));
    }
    float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (d + 1)); /* :179 */
    for (int i = 0; i < d; i++)                             /* :181-182 */
        scale_factor[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);

    for (int k = 0; k < N; k += B) { /* :191 */
        for (int j = 0; j < d; j++)  /* :193-196 */
            for (int i = 0; i < B; i++)
                f[j * B + i] = k + i < N ? feature[(size_t)(k + i) * d + j] : 0.0f;
        for (int l = 0; l < B; l++) {
            /* elevate :201-207 */
            float sm = 0.0f;
            for (int j = d; j > 0; j--) {
                float cf = f[(j - 1) * B + l] * scale_factor[j - 1];
                float jc = (float)j * cf;
                elevated[j * B + l] = sm - jc;
                sm = sm + cf;
            }
            elevated[0 * B + l] = sm;
            /* closest 0-coloured simplex :210-220 (cvtps_epi32 = round half even) */
            float sum = 0.0f;
            for (int i = 0; i <= d; i++) {
                float v = invdplus1 * elevated[i * B + l];
                v = (float)lrintf(v);
                rem0[i * B + l] = v * dplus1;
                sum = sum + v;
            }
            /* rank :223-233 */
            for (int i = 0; i <= d; i++) rank[i * B + l] = 0.0f;
            for (int i = 0; i < d; i++) {
                float di = elevated[i * B + l] - rem0[i * B + l];
                for (int j = i + 1; j <= d; j++) {
                    float dj = elevated[j * B + l] - rem0[j * B + l];
                    float c = di < dj ? 1.0f : 0.0f;
                    rank[i * B + l] += c;
                    rank[j * B + l] += 1.0f - c;
                }
            }
            /* bring back to the plane :236-242 */
            for (int i = 0; i <= d; i++) {
                rank[i * B + l] += sum;
                float add = rank[i * B + l] < 0.0f ? dplus1 : 0.0f;
                float sub = rank[i * B + l] >= dplus1 ? dplus1 : 0.0f;
                rank[i * B + l] += add - sub;
                rem0[i * B + l] += add - sub;
            }
        }
        /* barycentric :245-258 */
        for (int i = 0; i < (d + 2) * B; i++) barycentric[i] = 0.0f;
        for (int i = 0; i <= d; i++)
            for (int j = 0; j < B; j++) {
                float v = (elevated[i * B + j] - rem0[i * B + j]) * invdplus1;
                int p = (int)((float)d - rank[i * B + j]);
                barycentric[j * (d + 2) + p] += v;
                barycentric[j * (d + 2) + p + 1] -= v;
            }
        /* vertices :261-276 */
        for (int j = 0; j < B; j++) {
            barycentric[j * (d + 2) + 0] += 1 + barycentric[j * (d + 2) + d + 1];
            for (int remainder = 0; remainder <= d; remainder++) {
                for (int i = 0; i < d; i++)
                    key[i] = (short)(rem0[i * B + j] +
                                     canonical[remainder * (d + 1) + (int)rank[i * B + j]]);
                size_t at = (size_t)(j + k) * (d + 1) + remainder;
                L->offset[at] = ohash_find(&ht, key, 1);
                L->rank[at] = (int)rank[remainder * B + j];
                L->bary[at] = barycentric[j * (d + 2) + remainder];
            }
        }
    }
    free(scale_factor);
    free(f);
    free(elevated);
    free(rem0);
    free(rank);
    free(barycentric);
    free(canonical);
    free(key);

    /* blur neighbours :296-318 */
    L->M = (int)ht.filled;
    int M = L->M;
    L->n1 = (int *)malloc(sizeof(int) * (size_t)(d + 1) * M);
    L->n2 = (int *)malloc(sizeof(int) * (size_t)(d + 1) * M);
    short *n1 = (short *)malloc(sizeof(short) * (d + 1));
    short *n2 = (short *)malloc(sizeof(short) * (d + 1));
    for (int j = 0; j <= d; j++)
        for (int i = 0; i < M; i++) {
            const short *kk = ht.keys + (size_t)i * d;
            for (int k = 0; k < d; k++) {
                n1[k] = (short)(kk[k] - 1);
                n2[k] = (short)(kk[k] + 1);
            }
            /* for j == d this reads/writes slot d, outside the d hashed coords */
            n1[j] = (short)((j < d ? kk[j] : 0) + d);
            n2[j] = (short)((j < d ? kk[j] : 0) - d);
            L->n1[(size_t)j * M + i] = ohash_find(&ht, n1, 0);
            L->n2[(size_t)j * M + i] = ohash_find(&ht, n2, 0);
        }
    free(n1);
    free(n2);
    ohash_free(&ht);
    return L;
}

int oracle_lattice_M(const OLattice *L) { return L->M; }
const int *oracle_lattice_offset(const OLattice *L) { return L->offset; }
const float *oracle_lattice_bary(const OLattice *L) { return L->bary; }
const int *oracle_lattice_rank(const OLattice *L) { return L->rank; }
const int *oracle_lattice_n1(const OLattice *L) { return L->n1; }
const int *oracle_lattice_n2(const OLattice *L) { return L->n2; }

/* Permutohedral::seqCompute, CRF/src/permutohedral.cpp:476-527 (forward order only:
 * DenseKernel::filter is always called with transpose=false on this path). */
void oracle_lattice_seq_compute(const OLattice *L, float *out, const float *in, int vs) {
    int N = L->N, M = L->M, d = L->d;
    size_t tot = (size_t)(M + 2) * vs;
    float *values = (float *)calloc(tot, sizeof(float));
    float *new_values = (float *)calloc(tot, sizeof(float));
    for (int i = 0; i < N; i++) /* splat :486-493 */
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * (d + 1) + j] + 1;
            float w = L->bary[(size_t)i * (d + 1) + j];
            for (int k = 0; k < vs; k++) values[(size_t)o * vs + k] += w * in[(size_t)i * vs + k];
        }
    for (int j = 0; j <= d; j++) { /* blur :495-508, the add is done in double (:505) */
        for (int i = 0; i < M; i++) {
            float *old_val = values + (size_t)(i + 1) * vs;
            float *new_val = new_values + (size_t)(i + 1) * vs;
            int n1 = L->n1[(size_t)j * M + i] + 1;
            int n2 = L->n2[(size_t)j * M + i] + 1;
            float *n1_val = values + (size_t)n1 * vs;
            float *n2_val = values + (size_t)n2 * vs;
            for (int k = 0; k < vs; k++)
                new_val[k] = (float)((double)old_val[k] + 0.5 * (double)(n1_val[k] + n2_val[k]));
        }
        float *t = values;
        values = new_values;
        new_values = t;
    }
    float alpha = 1.0f / (1 + powf(2, -d)); /* :510 */
    for (int i = 0; i < N; i++) {           /* slice :513-522 */
        for (int k = 0; k < vs; k++) out[(size_t)i * vs + k] = 0;
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * (d + 1) + j] + 1;
            float w = L->bary[(size_t)i * (d + 1) + j];
            for (int k = 0; k < vs; k++)
                out[(size_t)i * vs + k] += w * values[(size_t)o * vs + k] * alpha;
        }
    }
    free(values);
    free(new_values);
}

/* Permutohedral::sseCompute, CRF/src/permutohedral.cpp:529-589.  The value dimension
 * is padded to a multiple of 4 there; padding lanes stay zero and every lane is an
 * independent IEEE-single op, so a per-channel scalar loop is arithmetically the
 * same.  in/out may alias (the reference calls compute(out, out), pairwise.cpp:74). */
void oracle_lattice_sse_compute(const OLattice *L, float *out, const float *in, int vs) {
    int N = L->N, M = L->M, d = L->d;
    size_t tot = (size_t)(M + 2) * vs;
    float *values = (float *)calloc(tot, sizeof(float));
    float *new_values = (float *)calloc(tot, sizeof(float));
    for (int i = 0; i < N; i++) /* splat :545-553 */
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * (d + 1) + j] + 1;
            float w = L->bary[(size_t)i * (d + 1) + j];
            for (int k = 0; k < vs; k++) values[(size_t)o * vs + k] += w * in[(size_t)i * vs + k];
        }
    for (int j = 0; j <= d; j++) { /* blur :555-569 */
        for (int i = 0; i < M; i++) {
            float *old_val = values + (size_t)(i + 1) * vs;
            float *new_val = new_values + (size_t)(i + 1) * vs;
            int n1 = L->n1[(size_t)j * M + i] + 1;
            int n2 = L->n2[(size_t)j * M + i] + 1;
            float *n1_val = values + (size_t)n1 * vs;
            float *n2_val = values + (size_t)n2 * vs;
            for (int k = 0; k < vs; k++) {
                float s = n1_val[k] + n2_val[k];
                float hs = 0.5f * s;
                new_val[k] = old_val[k] + hs;
            }
        }
        float *t = values;
        values = new_values;
        new_values = t;
    }
    float alpha = 1.0f / (1 + powf(2, -d)); /* :571 */
    float *acc = (float *)malloc(sizeof(float) * vs);
    for (int i = 0; i < N; i++) { /* slice :574-584 */
        for (int k = 0; k < vs; k++) acc[k] = 0;
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * (d + 1) + j] + 1;
            float w = L->bary[(size_t)i * (d + 1) + j] * alpha;
            for (int k = 0; k < vs; k++) acc[k] += w * values[(size_t)o * vs + k];
        }
        memcpy(out + (size_t)i * vs, acc, sizeof(float) * vs);
    }
    free(acc);
    free(values);
    free(new_values);
}

/* Permutohedral::compute dispatch, CRF/src/permutohedral.cpp:596-604 */
static void lattice_compute(const OLattice *L, float *out, const float *in, int vs) {
    if (vs <= 2)
        oracle_lattice_seq_compute(L, out, in, vs);
    else
        oracle_lattice_sse_compute(L, out, in, vs);
}

/* ------------------------------------------------------------------ */
/* Dense CRF                                                            */
/* ------------------------------------------------------------------ */
typedef struct {
    OLattice *lat;
    float *norm; /* N */
    float w;     /* Potts weight */
} OPairwise;

/* DenseKernel::initLattice, NORMALIZE_SYMMETRIC branch: CRF/src/pairwise.cpp:40-62 */
static void pairwise_init(OPairwise *P, const float *feature, int d, int N, float w) {
    P->lat = oracle_lattice_init(feature, d, N);
    P->w = w;
    P->norm = (float *)malloc(sizeof(float) * N);
    float *ones = (float *)malloc(sizeof(float) * N);
    for (int i = 0; i < N; i++) ones[i] = 1.0f;
    lattice_compute(P->lat, P->norm, ones, 1); /* :44 */
    for (int i = 0; i < N; i++)                /* :55-56 */
        P->norm[i] = (float)(1.0 / sqrt((double)P->norm[i] + 1e-20));
    free(ones);
}

/* PairwisePotential::apply = DenseKernel::filter (pairwise.cpp:63-80, :173-178)
 * followed by PottsCompatibility::apply out = -w*Q (labelcompatibility.cpp:46-48).
 * Q, out: M x N column-major == pixel-major [N][M]. */
static void pairwise_apply(const OPairwise *P, float *out, const float *Q, int M, int N) {
    for (int i = 0; i < N; i++) /* :66 */
        for (int k = 0; k < M; k++) out[(size_t)i * M + k] = Q[(size_t)i * M + k] * P->norm[i];
    lattice_compute(P->lat, out, out, M); /* :74 */
    for (int i = 0; i < N; i++)           /* :79 then Potts */
        for (int k = 0; k < M; k++) {
            float v = out[(size_t)i * M + k] * P->norm[i];
            out[(size_t)i * M + k] = -P->w * v;
        }
}

/* expAndNormalize, CRF/src/densecrf.cpp:98-106 (Eigen exp/sum -> expf/sequential) */
static void exp_and_normalize(float *out, const float *in, int M, int N) {
    for (int i = 0; i < N; i++) {
        const float *b = in + (size_t)i * M;
        float *o = out + (size_t)i * M;
        float mx = b[0];
        for (int k = 1; k < M; k++) mx = b[k] > mx ? b[k] : mx;
        float s = 0.0f;
        for (int k = 0; k < M; k++) {
            o[k] = expf(b[k] - mx);
            s += o[k];
        }
        for (int k = 0; k < M; k++) o[k] = o[k] / s;
    }
}

typedef struct {
    int W, H, M, N;
    float *unary; /* [N][M] energies */
    int n_pairwise;
    OPairwise pw[2];
} OCRF;

/* DenseCRFWrapper ctor, CRF/src/densecrf_wrapper.cpp:5-8 */
OCRF *oracle_crf_create(int W, int H, int M) {
    OCRF *c = (OCRF *)calloc(1, sizeof(OCRF));
    c->W = W;
    c->H = H;
    c->M = M;
    c->N = W * H;
    c->unary = (float *)calloc((size_t)c->N * M, sizeof(float)); /* densecrf.cpp:117 */
    return c;
}

void oracle_crf_destroy(OCRF *c) {
    if (!c) return;
    for (int k = 0; k < c->n_pairwise; k++) {
        oracle_lattice_free(c->pw[k].lat);
        free(c->pw[k].norm);
    }
    free(c->unary);
    free(c);
}

/* DenseCRFWrapper::set_unary_energy, densecrf_wrapper.cpp:32-37 (col-major M x N map) */
void oracle_crf_set_unary_energy(OCRF *c, const float *unary_costs) {
    memcpy(c->unary, unary_costs, sizeof(float) * (size_t)c->N * c->M);
}

/* DenseCRFWrapper::add_pairwise_energy, densecrf_wrapper.cpp:18-30: Gaussian (w2)
 * FIRST, then bilateral (w1); features from densecrf.cpp:61-69 and :70-81. */
void oracle_crf_add_pairwise_energy(OCRF *c, float w1, float ta1, float ta2, float tb1, float tb2,
                                    float tb3, float w2, float tg1, float tg2,
                                    const unsigned char *im) {
    int W = c->W, H = c->H, N = c->N;
    float *f2 = (float *)calloc(2 * (size_t)N, sizeof(float));
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            f2[(size_t)(j * W + i) * 2 + 0] = i / tg1;
            f2[(size_t)(j * W + i) * 2 + 1] = j / tg2;
        }
    pairwise_init(&c->pw[0], f2, 2, N, w2);
    free(f2);
    float *f5 = (float *)calloc(5 * (size_t)N, sizeof(float));
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            size_t p = (size_t)(j * W + i);
            f5[p * 5 + 0] = i / ta1;
            f5[p * 5 + 1] = j / ta2;
            f5[p * 5 + 2] = im[(i + j * W) * 3 + 0] / tb1;
            f5[p * 5 + 3] = im[(i + j * W) * 3 + 1] / tb2;
            f5[p * 5 + 4] = im[(i + j * W) * 3 + 2] / tb3;
        }
    pairwise_init(&c->pw[1], f5, 5, N, w1);
    free(f5);
    c->n_pairwise = 2;
}

/* DenseCRF::inference (densecrf.cpp:115-131) + DenseCRFWrapper::inference
 * (densecrf_wrapper.cpp:45-50; probs_out is pixel-major [N][M]). */
void oracle_crf_inference(OCRF *c, int n_iters, float *probs_out) {
    int M = c->M, N = c->N;
    size_t sz = (size_t)N * M;
    float *Q = (float *)malloc(sizeof(float) * sz);
    float *tmp1 = (float *)malloc(sizeof(float) * sz);
    float *tmp2 = (float *)malloc(sizeof(float) * sz);
    for (size_t i = 0; i < sz; i++) tmp1[i] = -c->unary[i];
    exp_and_normalize(Q, tmp1, M, N); /* :120 */
    for (int it = 0; it < n_iters; it++) {
        for (size_t i = 0; i < sz; i++) tmp1[i] = -c->unary[i]; /* :123 */
        for (int k = 0; k < c->n_pairwise; k++) {
            pairwise_apply(&c->pw[k], tmp2, Q, M, N);         /* :125 */
            for (size_t i = 0; i < sz; i++) tmp1[i] -= tmp2[i]; /* :126 */
        }
        exp_and_normalize(Q, tmp1, M, N); /* :128 */
    }
    memcpy(probs_out, Q, sizeof(float) * sz);
    free(Q);
    free(tmp1);
    free(tmp2);
}

/* DenseCRF::map + currentMap (densecrf.cpp:132-137, :202-211): first maximum wins. */
void oracle_crf_map(OCRF *c, int n_iters, int *labels) {
    int M = c->M, N = c->N;
    float *Q = (float *)malloc(sizeof(float) * (size_t)N * M);
    oracle_crf_inference(c, n_iters, Q);
    for (int i = 0; i < N; i++) {
        int m = 0;
        for (int k = 1; k < M; k++)
            if (Q[(size_t)i * M + k] > Q[(size_t)i * M + m]) m = k;
        labels[i] = m;
    }
    free(Q);
}

/* introspection for lattice-level tests */
const OLattice *oracle_crf_lattice(const OCRF *c, int k) { return c->pw[k].lat; }
const float *oracle_crf_norm(const OCRF *c, int k) { return c->pw[k].norm; }
