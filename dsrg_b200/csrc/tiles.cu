// Tile-local lattice views (sm_100a).
//
// The permutohedral splat / slice of the reference touch, per pixel, d+1 lattice rows of 24
// floats each (CRF/src/permutohedral.cpp:545-553, :574-584).  Neighbouring pixels share most of
// their vertices, so a 32x8-pixel tile only touches a few dozen DISTINCT rows.  This kernel
// finds them once per lattice build: per tile the distinct rows with, for each, its segment of the
// transposed incidence (tl_hdr), per pixel the index into that list (tl_loc), and the incidence
// itself as a CSR grouped by local vertex whose entries are already in the form the mean-field
// kernel consumes (tl_pack: byte offset of the pixel's Q row inside the tile, weight).  Segments
// are padded to multiples of four entries (zero weight) so that the consumer's loop has no
// remainder and the whole entry list is one 16-byte-aligned block for a bulk copy.
// The symmetric normalisation is folded into the weights: wn = bary * norm (pairwise.cpp:66,79).
#include "common.cuh"

namespace dsrg {

template <int DP1, int MAXLOC, int MP>
__global__ void __launch_bounds__(256)
k_tile_build(const int32_t *off, const float *bary, const float *norm, int32_t *tl_nloc, int2 *tl_hdr,
             int2 *tl_pack, uint16_t *tl_loc, float *wn, int N, int W, int H, int tiles_x, int ntiles,
             int entcap) {
    constexpr int HS = 2048;  // >= 256*DP1 distinct rows in the worst case, power of two
    static_assert(MAXLOC <= 256, "one scan element per thread");
    __shared__ int hkey[HS];
    __shared__ int hlv[HS];
    __shared__ int rows_s[MAXLOC];
    __shared__ int cnt[MAXLOC];
    __shared__ int ptr[MAXLOC + 1];
    __shared__ int wsum[8];
    __shared__ int count;

    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int x = tx * kTileW + (tid & 31), y = ty * kTileH + (tid >> 5);
    const bool in = x < W && y < H;
    const int pix = y * W + x;
    for (int i = tid; i < HS; i += 256) hkey[i] = -1;
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
            unsigned s = ((unsigned)row * 2654435761u) >> 21;  // 11 bits
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
    int lv[DP1];
    if (in) {
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            lv[r] = hlv[slot[r]];
            tl_loc[((size_t)b * DP1 + r) * N + pix] = (uint16_t)lv[r];
            atomicAdd(&cnt[lv[r]], 1);
        }
    }
    __syncthreads();
    // exclusive scan of the PADDED segment lengths -> ptr
    const int mycnt = tid < MAXLOC ? cnt[tid] : 0;
    const int mypad = (mycnt + 3) & ~3;
    {
        int incl = mypad;
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
            ptr[tid] = base + incl - mypad;
            cnt[tid] = 0;  // reused as the fill cursor
        }
    }
    __syncthreads();
    int2 *pack = tl_pack + ((size_t)b * ntiles + tile) * entcap;
    if (in) {
#pragma unroll
        for (int r = 0; r < DP1; r++) {
            int pos = ptr[lv[r]] + atomicAdd(&cnt[lv[r]], 1);
            pack[pos] = make_int2(tid * (MP * 4), __float_as_int(w[r]));
        }
    }
    if (tid < nloc) {
        for (int k = mycnt; k < mypad; k++) pack[ptr[tid] + k] = make_int2(0, 0);  // zero-weight padding
        tl_hdr[((size_t)b * ntiles + tile) * MAXLOC + tid] = make_int2(ptr[tid] | ((mypad >> 2) << 16), rows_s[tid]);
    }
    if (tid == 0) tl_nloc[tile] = nloc;
}

int tiles_build(Engine *e, Lattice &L, int nb, cudaStream_t s) {
    dim3 g(e->ntiles, nb);
#define DSRG_TILE_BUILD(DP1, MAXLOC, MPV)                                                                   \
    DSRG_LAUNCH(e, T_LAT_MISC, s,                                                                            \
                (k_tile_build<DP1, MAXLOC, MPV><<<g, 256, 0, s>>>(L.off, L.bary, L.norm, L.tl_nloc, L.tl_hdr, \
                                                                  L.tl_pack, L.tl_loc, L.wn, L.N, e->W, e->H,  \
                                                                  e->tiles_x, e->ntiles, L.entcap)))
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
