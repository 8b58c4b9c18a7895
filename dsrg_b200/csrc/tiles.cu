// Tile-local lattice views (sm_100a).
//
// The permutohedral splat / slice of the reference touch, per pixel, d+1 lattice rows of 24
// floats each (CRF/src/permutohedral.cpp:545-553, :574-584).  Neighbouring pixels share most of
// their vertices, so a 32x8-pixel tile only touches a few dozen DISTINCT rows.  This kernel
// finds them once per lattice build: per tile the distinct rows with, for each, its segment of the
// transposed incidence (tl_hdr), per pixel the index into that list (tl_loc), and the incidence
// itself as a CSR grouped by local vertex whose entries are already in the form the mean-field
// kernel consumes (tl_pack: byte offset of the pixel's Q row inside the tile, weight); the whole
// entry list of a tile is one 16-byte-aligned block for a bulk copy.  Local vertices are numbered
// by decreasing segment length.
// The symmetric normalisation is folded into the weights: wn = bary * norm (pairwise.cpp:66,79).
#include "common.cuh"

namespace dsrg {

template <int DP1, int MAXLOC, int MP>
__global__ void __launch_bounds__(kTileThreads)
k_tile_build(const int32_t *off, const float *bary, const float *norm, int32_t *tl_nloc, int2 *tl_hdr,
             int2 *tl_pack, uint16_t *tl_loc, float *wn, int N, int W, int H, int tiles_x, int ntiles,
             int entcap, int tile_w) {
    constexpr int HS = kTileThreads * 8;  // >= kTileThreads*DP1 distinct rows in the worst case, power of two
    constexpr int HBITS = (HS == 1024) ? 10 : (HS == 2048) ? 11 : (HS == 4096) ? 12 : -1;
    static_assert(HBITS > 0, "hash size");
    static_assert(MAXLOC <= kTileThreads, "one scan element per thread");
    __shared__ int hkey[HS];
    __shared__ int hlv[HS];
    __shared__ int rows_s[MAXLOC];
    __shared__ int cnt[MAXLOC];
    __shared__ int ptr[MAXLOC + 1];
    __shared__ int wsum[kTileThreads / 32];
    __shared__ int perm[MAXLOC];
    __shared__ int scnt[MAXLOC];
    __shared__ int count;

    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int x = tx * tile_w + (tid & 31), y = ty * kTileH + (tid >> 5);
    const bool in = (tid & 31) < tile_w && x < W && y < H;
    const int pix = y * W + x;
    for (int i = tid; i < HS; i += kTileThreads) hkey[i] = -1;
    if (tid < MAXLOC) cnt[tid] = 0;
    if (tid == 0) count = 0;
    __syncthreads();

    int slot[DP1];
    float w[DP1];
    if (in) {
        const float nrm = norm[(size_t)b * N + pix];
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            const size_t at = ((size_t)b * DP1 + r) * N + pix;
            const int row = off[at];
            w[r] = __fmul_rn(bary[at], nrm);
            wn[at] = w[r];
            unsigned s = ((unsigned)row * 2654435761u) >> (32 - HBITS);
            while (true) {
                int old = atomicCAS(&hkey[s], -1, row);
                if (old == -1) {
                    int lv = atomicAdd(&count, 1);
                    hlv[s] = lv;
                    if (lv < MAXLOC) rows_s[lv] = row;
                    break;
                }
                if (old == row) break;
                s = (s + 1) & (HS - 1);
            }
            slot[r] = (int)s;
        }
    }
    __syncthreads();
    const int nloc = count;
    tl_nloc += (size_t)b * ntiles;
    if (nloc > MAXLOC) {  // too many distinct vertices for the shared-memory path: fallback tile
        if (tid == 0) tl_nloc[tile] = -1;
        return;
    }
    // order the local vertices by decreasing segment length so that the lanes of a warp of the
    // consumer (one lane per (vertex, label quad)) walk segments of similar length
    int lv[DP1];
    if (in) {
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            lv[r] = hlv[slot[r]];
            atomicAdd(&cnt[lv[r]], 1);
        }
    }
    __syncthreads();
    const int mycnt = tid < nloc ? cnt[tid] : 0;
    if (tid < nloc) {
        int rank = 0;
        for (int j = 0; j < nloc; j++) {
            const int cj = cnt[j];
            rank += (cj > mycnt) || (cj == mycnt && j < tid);
        }
        perm[tid] = rank;      // new index of local vertex `tid`
        scnt[rank] = mycnt;    // counts in the new order
    }
    __syncthreads();
    // exclusive scan of the reordered counts, each rounded up to an even number -> ptr: every segment then
    // starts on a 16-byte boundary and the consumer reads its entries two at a time (one LDS.128)
    {
        const int v = tid < nloc ? ((scnt[tid] + 1) & ~1) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += n;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        for (int k = 0; k < (tid >> 5); k++) base += wsum[k];
        if (tid < MAXLOC) {
            ptr[tid] = base + incl - v;
            cnt[tid] = 0;  // reused as the fill cursor (indexed by the NEW vertex index)
        }
    }
    __syncthreads();
    int2 *pack = tl_pack + ((size_t)b * ntiles + tile) * entcap;
    if (in) {
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            const int nv = perm[lv[r]];
            tl_loc[((size_t)b * DP1 + r) * N + pix] = (uint16_t)nv;
            const int pos = ptr[nv] + atomicAdd(&cnt[nv], 1);
            pack[pos] = make_int2(tid * ((MP + 4 * DSRG_ROW_PAD) * 4), __float_as_int(w[r]));  // byte offset of the pixel's (padded) Q row
        }
    }
    if (tid < nloc) {
        const int nv = perm[tid];
        tl_hdr[((size_t)b * ntiles + tile) * MAXLOC + nv] = make_int2(ptr[nv] | (mycnt << 16), rows_s[tid]);
        if (mycnt & 1) pack[ptr[nv] + mycnt] = make_int2(0, 0);  // padding entry: pixel 0 with weight 0
    }
    if (tid == 0) tl_nloc[tile] = nloc;
}

int tiles_build(Engine *e, Lattice &L, int nb, cudaStream_t s) {
    dim3 g(e->ntiles, nb);
#define DSRG_TILE_BUILD(DP1, MAXLOC, MPV)                                                                   \
    DSRG_LAUNCH(e, T_LAT_MISC, s,                                                                            \
                (k_tile_build<DP1, MAXLOC, MPV><<<g, kTileThreads, 0, s>>>(L.off, L.bary, L.norm, L.tl_nloc, L.tl_hdr, \
                                                                  L.tl_pack, L.tl_loc, L.wn, L.N, e->W, e->H,  \
                                                                  e->tiles_x, e->ntiles, L.entcap, e->tile_w)))
#define DSRG_TILE_BUILD_MP(MPV)                           \
    if (L.d == 2) { DSRG_TILE_BUILD(3, kMaxLocSp, MPV); } \
    else { DSRG_TILE_BUILD(6, kMaxLocBi, MPV); }
    switch (e->MP) {
        case 4: DSRG_TILE_BUILD_MP(4); break;
        case 8: DSRG_TILE_BUILD_MP(8); break;
        case 12: DSRG_TILE_BUILD_MP(12); break;
        case 16: DSRG_TILE_BUILD_MP(16); break;
        case 20: DSRG_TILE_BUILD_MP(20); break;
        case 24: DSRG_TILE_BUILD_MP(24); break;
        case 28: DSRG_TILE_BUILD_MP(28); break;
        case 32: DSRG_TILE_BUILD_MP(32); break;
        default: set_error("unsupported label count"); return DSRG_E_INVALID;
    }
    DSRG_CUDA_TRY(cudaGetLastError());
    return DSRG_OK;
}

}  // namespace dsrg
