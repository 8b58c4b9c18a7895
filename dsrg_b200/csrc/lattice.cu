// Permutohedral lattice construction on the GPU (sm_100a).
//
// Replaces Permutohedral::init + HashTable (CRF/src/permutohedral.cpp:54-131, :140-321) and
// DenseKernel::initLattice (CRF/src/pairwise.cpp:40-62) of the reference, for a whole batch at
// once.  The arithmetic that decides WHICH simplex a pixel falls into follows the reference's
// SSE code path operation by operation (round-to-nearest-even, no FMA contraction: every float
// op is an explicit __f*_rn intrinsic); the hash table itself is a GPU design (64-bit packed
// keys, CAS insertion, ids in arrival order) because vertex numbering does not influence the
// filter result.
#include "common.cuh"

namespace dsrg {

// ---------------------------------------------------------------------------------------------
// packed keys: d=2 -> 2 x 16 bit (exactly the reference's `short`), d=5 -> 5 x 12 bit
// ---------------------------------------------------------------------------------------------
template <int D>
struct KeyBits {
    static constexpr int bits = (D == 2) ? 16 : 12;
    static constexpr int lo = -(1 << (bits - 1));
    static constexpr int hi = (1 << (bits - 1)) - 1;
};

template <int D>
__device__ __forceinline__ uint64_t pack_key(const int *key) {
    constexpr int BITS = KeyBits<D>::bits;
    uint64_t k = 0;
#pragma unroll
    for (int i = 0; i < D; i++) k |= (uint64_t)((uint32_t)key[i] & ((1u << BITS) - 1)) << (i * BITS);
    return k;
}

template <int D>
__device__ __forceinline__ void unpack_key(uint64_t k, int *key) {
    constexpr int BITS = KeyBits<D>::bits;
#pragma unroll
    for (int i = 0; i < D; i++) {
        int v = (int)((k >> (i * BITS)) & ((1u << BITS) - 1));
        key[i] = (v << (32 - BITS)) >> (32 - BITS);  // sign extend
    }
}

__device__ __forceinline__ uint32_t hash_slot(uint64_t k, uint32_t cap) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)(((k >> 32) * (uint64_t)cap) >> 32);
}

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// insert-or-find; the winner of an empty slot allocates the next vertex id of its image
__device__ __forceinline__ int hash_insert(uint64_t *keys, int32_t *hval, int32_t *vslot,
                                           int32_t *vcount, uint32_t cap, uint64_t k) {
    uint32_t s = hash_slot(k, cap);
    while (true) {
        uint64_t cur = ld_relaxed_u64(keys + s);
        if (cur == k) return (int)s;
        if (cur == kEmptyKey) {
            unsigned long long old =
                atomicCAS((unsigned long long *)(keys + s), (unsigned long long)kEmptyKey,
                          (unsigned long long)k);
            if (old == kEmptyKey) {
                int id = atomicAdd(vcount, 1);
                vslot[id] = (int)s;
                asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(hval + s), "r"(id) : "memory");
                return (int)s;
            }
            if (old == k) return (int)s;
        }
        s = (s + 1 == cap) ? 0 : s + 1;
    }
}

__device__ __forceinline__ int hash_lookup(const uint64_t *keys, uint32_t cap, uint64_t k) {
    uint32_t s = hash_slot(k, cap);
    while (true) {
        uint64_t cur = keys[s];
        if (cur == k) return (int)s;
        if (cur == kEmptyKey) return -1;
        s = (s + 1 == cap) ? 0 : s + 1;
    }
}

struct BuildArgs {
    int N, P, W;
    uint32_t cap;
    int capv;
    float sigma[5];
    float scale[5];
    const uint8_t *image;  // [B][N][3] or nullptr
    int32_t *off;
    float *bary;
    uint64_t *hkeys;
    int32_t *hval, *vslot, *vcount;
    int *err;
};

// ---------------------------------------------------------------------------------------------
// Kernel 1: one thread per pixel (plus the phantom tail lanes): features -> elevate -> simplex ->
// rank -> barycentric -> d+1 vertex keys -> hash insert.  Follows permutohedral.cpp:191-276.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads) k_lattice_insert(BuildArgs a) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.N + a.P;
    const bool real = i < a.N;

    float f[D];
#pragma unroll
    for (int j = 0; j < D; j++) f[j] = 0.0f;  // phantom lanes carry feature 0 (:196)
    if (real) {
        const int x = i % a.W, y = i / a.W;
        f[0] = __fdiv_rn((float)x, a.sigma[0]);  // densecrf.cpp:65-66 / :74-75
        f[1] = __fdiv_rn((float)y, a.sigma[1]);
        if (D == 5) {
            const uint8_t *px = a.image + ((size_t)b * a.N + i) * 3;  // densecrf.cpp:76-78
            f[2] = __fdiv_rn((float)px[0], a.sigma[2]);
            f[3] = __fdiv_rn((float)px[1], a.sigma[3]);
            f[4] = __fdiv_rn((float)px[2], a.sigma[4]);
        }
    }
    // elevate (:201-207)
    float el[D + 1];
    float sm = 0.0f;
#pragma unroll
    for (int j = D; j > 0; j--) {
        float cf = __fmul_rn(f[j - 1], a.scale[j - 1]);
        el[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
        sm = __fadd_rn(sm, cf);
    }
    el[0] = sm;
    // closest 0-coloured simplex (:210-220); cvtps_epi32 under MXCSR-nearest == rintf
    const float inv = __fdiv_rn(1.0f, (float)(D + 1));
    const float dp1 = (float)(D + 1);
    float rem0[D + 1];
    int isum = 0;
#pragma unroll
    for (int k = 0; k <= D; k++) {
        float v = rintf(__fmul_rn(inv, el[k]));
        rem0[k] = __fmul_rn(v, dp1);
        isum += (int)v;
    }
    // rank (:223-233): strict <, ties increment the later coordinate
    int rank[D + 1];
#pragma unroll
    for (int k = 0; k <= D; k++) rank[k] = 0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        float di = __fsub_rn(el[k], rem0[k]);
#pragma unroll
        for (int j = k + 1; j <= D; j++) {
            float dj = __fsub_rn(el[j], rem0[j]);
            if (di < dj) rank[k]++; else rank[j]++;
        }
    }
    // back onto the plane (:236-242)
#pragma unroll
    for (int k = 0; k <= D; k++) {
        rank[k] += isum;
        if (rank[k] < 0) {
            rank[k] += D + 1;
            rem0[k] = __fadd_rn(rem0[k], dp1);
        } else if (rank[k] >= D + 1) {
            rank[k] -= D + 1;
            rem0[k] = __fsub_rn(rem0[k], dp1);
        }
    }
    // barycentric (:245-263), same accumulation order as the reference
    float bc[D + 2];
#pragma unroll
    for (int q = 0; q <= D + 1; q++) bc[q] = 0.0f;
#pragma unroll
    for (int k = 0; k <= D; k++) {
        float v = __fmul_rn(__fsub_rn(el[k], rem0[k]), inv);
        int p = D - rank[k];
#pragma unroll
        for (int q = 0; q <= D + 1; q++) {
            if (q == p) bc[q] = __fadd_rn(bc[q], v);
            if (q == p + 1) bc[q] = __fsub_rn(bc[q], v);
        }
    }
    bc[0] = __fadd_rn(bc[0], __fadd_rn(1.0f, bc[D + 1]));

    // vertices (:268-275)
    uint64_t *keys = a.hkeys + (size_t)b * a.cap;
    int32_t *hval = a.hval + (size_t)b * a.cap;
    int32_t *vslot = a.vslot + (size_t)b * a.capv;
    int32_t *vcount = a.vcount + b;
    const unsigned lane = threadIdx.x & 31;
    bool range_bad = false;
#pragma unroll
    for (int r = 0; r <= D; r++) {
        int key[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            int canon = (rank[k] <= D - r) ? r : r - (D + 1);  // canonical simplex (:171-176)
            key[k] = (int)rem0[k] + canon;
            if (key[k] < KeyBits<D>::lo || key[k] > KeyBits<D>::hi) range_bad = true;
        }
        uint64_t pk = pack_key<D>(key);
        // neighbouring pixels mostly share vertices: one insert per distinct key per warp
        unsigned peers = __match_any_sync(0xffffffffu, valid ? pk : (kEmptyKey - 1 - lane));
        int leader = __ffs(peers) - 1;
        int id = 0;
        if (valid && (int)lane == leader) {
            const int slot = hash_insert(keys, hval, vslot, vcount, a.cap, pk);
            // the slot's owner publishes the vertex id right after winning the CAS; owners of the
            // same key always sit in other warps (one leader per key per warp), so this cannot
            // wait on a lane of the same warp
            do {
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(id) : "l"(hval + slot) : "memory");
            } while (id < 0);
        }
        id = __shfl_sync(0xffffffffu, id, leader);
        if (real) {
            size_t at = ((size_t)b * (D + 1) + r) * a.N + i;
            a.off[at] = id + 1;  // local row (row 0 of every image is its zero row)
            a.bary[at] = bc[r];
        }
    }
    if (valid && range_bad) *a.err = DSRG_E_KEYRANGE;
}

// rowbase[b] = first value row of image b (its zero row); rowbase[B] = total rows
__global__ void k_rowbase(const int32_t *vcount, int32_t *rowbase, int B, int shared) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b <= B; b++) {
            rowbase[b] = acc;
            if (b < B) acc += vcount[shared ? 0 : b] + 1;
        }
    }
}

// Kernel 3: blur neighbours of every vertex (:305-318).  Missing neighbour -> the zero row.
template <int D>
__global__ void __launch_bounds__(kThreads)
k_lattice_neighbors(const uint64_t *hkeys, const int32_t *hval, const int32_t *vslot,
                    const int32_t *vcount, const int32_t *rowbase, int2 *nbr, long long nbr_stride,
                    uint32_t cap, int capv, int shared) {
    const int b = blockIdx.y;
    const int V = vcount[b];
    const long long base = shared ? 0 : rowbase[b];
    const uint64_t *keys = hkeys + (size_t)b * cap;
    const int32_t *hv = hval + (size_t)b * cap;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= V; v += gridDim.x * blockDim.x) {
        if (v == V) {  // the zero row points at itself so that blurring keeps it at zero
#pragma unroll
            for (int j = 0; j <= D; j++) nbr[(size_t)j * nbr_stride + base] = make_int2((int)base, (int)base);
            continue;
        }
        int key[D];
        unpack_key<D>(keys[vslot[(size_t)b * capv + v]], key);
        const long long row = base + 1 + v;
#pragma unroll
        for (int j = 0; j <= D; j++) {
            int n1[D], n2[D];
            bool ok1 = true, ok2 = true;
#pragma unroll
            for (int k = 0; k < D; k++) {
                n1[k] = key[k] - 1;
                n2[k] = key[k] + 1;
                if (k == j) {
                    n1[k] = key[k] + D;
                    n2[k] = key[k] - D;
                }
                ok1 &= (n1[k] >= KeyBits<D>::lo && n1[k] <= KeyBits<D>::hi);
                ok2 &= (n2[k] >= KeyBits<D>::lo && n2[k] <= KeyBits<D>::hi);
            }
            int s1 = ok1 ? hash_lookup(keys, cap, pack_key<D>(n1)) : -1;
            int s2 = ok2 ? hash_lookup(keys, cap, pack_key<D>(n2)) : -1;
            int r1 = s1 < 0 ? (int)base : (int)(base + 1 + hv[s1]);
            int r2 = s2 < 0 ? (int)base : (int)(base + 1 + hv[s2]);
            nbr[(size_t)j * nbr_stride + row] = make_int2(r1, r2);
        }
    }
}

// Kernel 4: give the slots back (the tables stay all-empty between batches, no big memset)
__global__ void __launch_bounds__(kThreads)
k_lattice_cleanup(uint64_t *hkeys, int32_t *hval, const int32_t *vslot, const int32_t *vcount, uint32_t cap,
                  int capv) {
    const int b = blockIdx.y;
    const int V = vcount[b];
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const size_t at = (size_t)b * cap + vslot[(size_t)b * capv + v];
        hkeys[at] = kEmptyKey;
        hval[at] = -1;
    }
}

// ---------------------------------------------------------------------------------------------
// Normalisation: norm = 1/sqrt(K 1 + 1e-20), pairwise.cpp:44,54-57, with K 1 evaluated like
// Permutohedral::seqCompute(value_size=1) (permutohedral.cpp:476-527).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_norm_splat(const int32_t *off, const float *bary, const int32_t *rowbase, float *nv, int N,
             int dp1) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int base = rowbase[b];
    for (int r = 0; r < dp1; r++) {
        size_t at = ((size_t)b * dp1 + r) * N + i;
        atomicAdd(nv + base + off[at], bary[at]);  // values[o] += w * 1 (:491)
    }
}

__global__ void __launch_bounds__(kThreads)
k_norm_blur(const float *in, float *out, const int2 *nbr, const int32_t *rowbase, int B) {
    const int rows = rowbase[B];
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < rows; g += gridDim.x * blockDim.x) {
        int2 n = nbr[g];
        // (float)(old + 0.5*(n1+n2)) evaluated in double (:505) == this single-rounding float form
        out[g] = __fadd_rn(in[g], __fmul_rn(0.5f, __fadd_rn(in[n.x], in[n.y])));
    }
}

__global__ void __launch_bounds__(kThreads)
k_norm_slice(const int32_t *off, const float *bary, const int32_t *rowbase, const float *nv,
             float *norm, int N, int dp1, float alpha) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int base = rowbase[b];
    float acc = 0.0f;
    for (int r = 0; r < dp1; r++) {
        size_t at = ((size_t)b * dp1 + r) * N + i;
        float w = bary[at];
        // out += w * values[o] * alpha (:520)
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(w, nv[base + off[at]]), alpha));
    }
    norm[(size_t)b * N + i] = (float)(1.0 / sqrt((double)acc + 1e-20));  // pairwise.cpp:56
}

template <int D>
static int build_impl(Engine *e, Lattice &L, int nb, const uint8_t *image, cudaStream_t s) {
    BuildArgs a;
    a.N = L.N;
    a.P = L.P;
    a.W = e->W;
    a.cap = (uint32_t)L.cap;
    a.capv = L.capv;
    for (int i = 0; i < 5; i++) {
        a.sigma[i] = L.sigma[i];
        a.scale[i] = L.scale[i];
    }
    a.image = image;
    a.off = L.off;
    a.bary = L.bary;
    a.hkeys = L.hkeys;
    a.hval = L.hval;
    a.vslot = L.vslot;
    a.vcount = L.vcount;
    a.err = e->dev_err;
    DSRG_CUDA_TRY(cudaMemsetAsync(L.vcount, 0, sizeof(int32_t) * nb, s));
    dim3 gp(cdiv(L.N + L.P, kThreads), nb);
    DSRG_LAUNCH(e, T_LAT_INSERT, s, k_lattice_insert<D><<<gp, kThreads, 0, s>>>(a));
    DSRG_LAUNCH(e, T_LAT_MISC, s, k_rowbase<<<1, 32, 0, s>>>(L.vcount, L.rowbase, L.shared ? e->maxB : nb, L.shared));
    dim3 gv(2 * e->sm_count, nb);
    DSRG_LAUNCH(e, T_LAT_MISC, s,
                k_lattice_neighbors<D><<<gv, kThreads, 0, s>>>(L.hkeys, L.hval, L.vslot, L.vcount, L.rowbase,
                                                               L.nbr, L.nbr_stride, a.cap, L.capv, L.shared));
    DSRG_LAUNCH(e, T_LAT_MISC, s, k_lattice_cleanup<<<gv, kThreads, 0, s>>>(L.hkeys, L.hval, L.vslot, L.vcount, a.cap, L.capv));
    return DSRG_OK;
}

// Build the lattices of `B` images (or the single shared one), then their norm vectors.
int lattice_build(Engine *e, Lattice &L, int B, const uint8_t *image_dev, cudaStream_t s) {
    const int nb = L.shared ? 1 : B;
    int rc = (L.d == 2) ? build_impl<2>(e, L, nb, image_dev, s) : build_impl<5>(e, L, nb, image_dev, s);
    if (rc) return rc;
    // norm pass on nb structures; value rows of the norm pass use the per-structure packing,
    // which for the shared lattice is simply image 0's rows [0, V+1).
    const int dp1 = L.d + 1;
    const long long rows_nb = L.shared ? (long long)L.capv + 1 : L.rows_cap;
    DSRG_CUDA_TRY(cudaMemsetAsync(e->nvA, 0, sizeof(float) * rows_nb, s));
    dim3 gp(cdiv(L.N, kThreads), nb);
    DSRG_LAUNCH(e, T_LAT_NORM, s, k_norm_splat<<<gp, kThreads, 0, s>>>(L.off, L.bary, L.rowbase, e->nvA, L.N, dp1));
    float *src = e->nvA, *dst = e->nvB;
    for (int j = 0; j < dp1; j++) {
        DSRG_LAUNCH(e, T_LAT_NORM, s,
                    k_norm_blur<<<4 * e->sm_count, kThreads, 0, s>>>(src, dst, L.nbr + (size_t)j * L.nbr_stride,
                                                                      L.rowbase, nb));
        float *t = src;
        src = dst;
        dst = t;
    }
    const float alpha = 1.0f / (1 + powf(2, -L.d));  // permutohedral.cpp:510
    DSRG_LAUNCH(e, T_LAT_NORM, s, k_norm_slice<<<gp, kThreads, 0, s>>>(L.off, L.bary, L.rowbase, src, L.norm, L.N, dp1, alpha));
    DSRG_CUDA_TRY(cudaGetLastError());
    // the tile-local views serve the fused kernel (up to DSRG_MAX_LABELS labels); the wide path only needs bary * norm
    return e->MP > DSRG_MAX_LABELS ? wide_weights(e, L, nb, s) : tiles_build(e, L, nb, s);
}

}  // namespace dsrg
