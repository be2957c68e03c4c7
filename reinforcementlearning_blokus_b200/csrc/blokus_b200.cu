// blokus_b200.cu -- kernels and the C ABI of libblokus_b200.so (sm_100a only).
//
// One warp owns one game state (see movegen_core.cuh).  The streaming kernels are warp-per-item grid-stride
// loops, the playout kernels hand out games through a per-launch ticket counter; grids are sized to fill the
// 148 SMs at the kernel's occupancy.  Nothing here uses tensor cores (no dense contraction on this path) and
// nothing falls back to the CPU.  One process drives one GPU (bk_init(device) binds the library to it).
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "movegen_core.cuh"

#define BK_THREADS 256
#ifndef BK_COUNT_CTAS
#define BK_COUNT_CTAS 4      /* k_movegen_count: 64 registers -> 4 CTAs (32 warps) per SM */
#endif
#define BK_WARPS (BK_THREADS / 32)

// ---------------------------------------------------------------------------------------------
// error plumbing / lifecycle
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

// One context per CUDA device, created by bk_init(device) and selected by the CALLER'S current device at every entry
// point (the device whose stream and buffers the call names).  Nothing else in this library is mutable: the piece
// tables live in per-device constant memory written by bk_init, errors are thread-local.
#define BK_TICKETS 256
#define BK_MAX_DEVICES 64
struct BkCtx {
    int inited;
    int sms;
    unsigned int* tickets;      // per-launch work counters of the playout / search kernels (device, BK_TICKETS of them)
    unsigned int next;
};
static BkCtx g_ctx[BK_MAX_DEVICES];
static std::mutex g_init_mutex;

static int fail(int code, const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) return fail(BK_ERR_CUDA, #expr ": %s", cudaGetErrorString(e__)); \
    } while (0)
static BkCtx* current_ctx() {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= BK_MAX_DEVICES || !g_ctx[d].inited) return nullptr;
    return &g_ctx[d];
}
#define REQUIRE_INIT() \
    BkCtx* cx = current_ctx(); \
    do { if (!cx) return fail(BK_ERR_NOT_INIT, "%s", "bk_init has not been called for the current CUDA device"); } while (0)
#define REQUIRE(cond, msg) \
    do { if (!(cond)) return fail(BK_ERR_ARG, "%s", msg); } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// A zeroed work counter for one launch on `st` (round-robin over a small pool so that launches in flight on
// different streams never share one).
static cudaError_t fresh_ticket(BkCtx* cx, cudaStream_t st, unsigned int** out) {
    unsigned int* t = cx->tickets + (__sync_fetch_and_add(&cx->next, 1u) % BK_TICKETS);
    *out = t;
    return cudaMemsetAsync(t, 0, sizeof(unsigned int), st);
}

static inline int grid_for(const BkCtx* cx, int n_items, int ctas_per_sm) {
    long long need = ((long long)n_items + BK_WARPS - 1) / BK_WARPS;
    long long cap = (long long)cx->sms * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" const char* bk_last_error(void) { return g_err; }
extern "C" const char* bk_version(void) { return "blokus_b200 r2 sm_100a (row-sliced warp-per-game)"; }

__global__ void k_movegen_emit(const bk_state* __restrict__, const uint8_t* __restrict__, int, const uint32_t* __restrict__,
                               uint32_t, uint32_t* __restrict__, uint16_t* __restrict__);
__global__ void k_movegen_bitmap(const bk_state* __restrict__, const uint8_t* __restrict__, int, uint32_t* __restrict__);
#define BK_EMIT_SMEM_BYTES (8 * (32 * 60 * 4))

static cudaError_t bk_init_policies_impl();
static cudaError_t bk_init_policies() { return bk_init_policies_impl(); }
static cudaError_t bk_init_wide();

struct DeviceGuard {            // bk_init works on `device` and leaves the caller's current device as it found it
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" int bk_init(int device) {
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count || device >= BK_MAX_DEVICES) return fail(BK_ERR_ARG, "%s", "bk_init: no such CUDA device");
    std::lock_guard<std::mutex> lock(g_init_mutex);
    BkCtx* cx = &g_ctx[device];
    if (cx->inited) return BK_OK;
    DeviceGuard guard;
    if (cudaGetDevice(&guard.prev) != cudaSuccess) guard.prev = -1;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceGetAttribute(&cx->sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaMemcpyToSymbol(c_orient_piece, BK_TBL_ORIENT_PIECE, sizeof(BK_TBL_ORIENT_PIECE)));
    CUDA_TRY(cudaMemcpyToSymbol(c_orient_ncells, BK_TBL_ORIENT_NCELLS, sizeof(BK_TBL_ORIENT_NCELLS)));
    CUDA_TRY(cudaMemcpyToSymbol(c_orient_cells, BK_TBL_ORIENT_CELLS, sizeof(BK_TBL_ORIENT_CELLS)));
    CUDA_TRY(cudaMemcpyToSymbol(c_group_first, BK_TBL_GROUP_FIRST, sizeof(BK_TBL_GROUP_FIRST)));
    CUDA_TRY(cudaMemcpyToSymbol(c_piece_first, BK_TBL_PIECE_FIRST, sizeof(BK_TBL_PIECE_FIRST)));
    CUDA_TRY(cudaMemcpyToSymbol(c_orient_cells32, BK_TBL_ORIENT_CELLS32, sizeof(BK_TBL_ORIENT_CELLS32)));
    CUDA_TRY(cudaMemcpyToSymbol(c_orient_rowpat, BK_TBL_ORIENT_ROWPAT, sizeof(BK_TBL_ORIENT_ROWPAT)));
    CUDA_TRY(cudaFuncSetAttribute(k_movegen_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_EMIT_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_movegen_bitmap, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_EMIT_SMEM_BYTES));
    CUDA_TRY(bk_init_policies());
    CUDA_TRY(bk_init_wide());
    if (!cx->tickets) CUDA_TRY(cudaMalloc(&cx->tickets, BK_TICKETS * sizeof(unsigned int)));
    CUDA_TRY(cudaDeviceSynchronize());
    cx->inited = 1;
    return BK_OK;
}

// ---------------------------------------------------------------------------------------------
// K1: legal-move generation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BK_THREADS)
k_movegen_count(const bk_state* __restrict__ states, const uint8_t* __restrict__ players, int n,
                uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint64_t word = warp < n ? fetch_state_word(states + warp, lane) : 0ull;
    for (int i = warp; i < n; i += nwarps) {
        const uint64_t next = (i + nwarps < n) ? fetch_state_word(states + i + nwarps, lane) : 0ull;   // prefetch
        WarpState s; unpack_state(word, lane, s);
        word = next;
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        uint32_t A, C; make_masks(s, p, lane, A, C);
        CountSink sink; bk_gen_all<CountSink, true>(A, C, sel4(s.used, p), sink);
        const uint32_t tot = __reduce_add_sync(BK_FULL, sink.acc);
        if (lane == 0) counts[i] = tot;
    }
}

// Ascending move ids.  The legal-anchor rows of an orientation (word w = o*20 + row; bit `col` of word w is move id
// 20*w + col) are sparse: ~9 % of the 1 820 words of a mid-game state hold a move.  The sink therefore appends only
// the NON-EMPTY rows to a list in shared memory while the generator runs -- one ballot per orientation tells every
// lane where its row goes, and the list comes out in ascending id order.  Then the entries are dealt out evenly
// (lane j takes entries j*per .. (j+1)*per), one warp scan of the lanes' move counts gives their list positions and
// every lane writes the ids of its entries.  No zeroing, no pass over empty words, no search.  (A uniform branch around
// the append -- REDUX.OR of the rows first -- was measured slower: 0.33 vs 0.47 G generations/s; the compiler predicates
// the ten instructions of the append, which is cheaper than 91 more reductions.)
// (Round 1 staged all 1 820 words and located every list position with two binary searches: 1 885 warp instructions
// per state, LSU pipe 56 %; a first rewrite with contiguous 60-word slices per lane still spent 55 % of its 1 549
// instructions in lock-step bit loops over mostly empty words -- profiles/r02_emit_ncu.md.)
#define EMIT_WORDS_PER_LANE 60            /* 32 * 60 = 1920 >= 1820 (row staging of the bitmap generator) */
#define EMIT_ROW_WORDS (32 * EMIT_WORDS_PER_LANE)
#define EMIT_SMEM_PER_WARP (EMIT_ROW_WORDS * 4)
static_assert(BK_EMIT_SMEM_BYTES == BK_WARPS * EMIT_SMEM_PER_WARP, "emit shared-memory size");

struct SmemRowSink {
    static constexpr bool kRows = true;
    uint32_t* rows; int lane;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) {
        if (lane < 20) rows[O * 20 + lane] = L;
    }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
};

struct NonEmptyRowSink {               // list[k] = (word index << 20) | MIRRORED row bits (bit j = column 19 - j), ascending word index
    static constexpr bool kRows = false;  // rows as the mirrored generator makes them: no BREV per orientation, the extraction decodes
    uint32_t a0, a, lane20, below;        // a = 32-bit shared-memory address of the next free slot (see HeurPrivSink on why)
    __device__ __forceinline__ void begin(uint32_t* l, int lane_) {
        a0 = a = (uint32_t)__cvta_generic_to_shared(l); lane20 = (uint32_t)lane_ << 20; below = (1u << lane_) - 1u;
    }
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) {
        const uint32_t m = __ballot_sync(BK_FULL, L != 0u);
        if (m) {
            if (L) asm volatile("st.shared.u32 [%0], %1;" :: "r"(a + 4u * (uint32_t)__popc(m & below)), "r"(L + lane20 + ((uint32_t)(O * 20) << 20)) : "memory");
            a += 4u * (uint32_t)__popc(m);
        }
    }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
    __device__ __forceinline__ uint32_t count() const { return (a - a0) >> 2; }
};

#ifndef BK_EMIT_CTAS
#define BK_EMIT_CTAS 3
#endif
__global__ void __launch_bounds__(BK_THREADS, BK_EMIT_CTAS)
k_movegen_emit(const bk_state* __restrict__ states, const uint8_t* __restrict__ players, int n,
               const uint32_t* __restrict__ offsets, uint32_t cap, uint32_t* __restrict__ counts_out,
               uint16_t* __restrict__ ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* list = reinterpret_cast<uint32_t*>(smem_raw + (size_t)wib * EMIT_SMEM_PER_WARP);   // <= 1 820 entries
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint64_t word = warp < n ? fetch_state_word(states + warp, lane) : 0ull;
    for (int i = warp; i < n; i += nwarps) {
        const uint64_t next = (i + nwarps < n) ? fetch_state_word(states + i + nwarps, lane) : 0ull;   // prefetch
        WarpState s; unpack_state(word, lane, s);
        word = next;
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        uint32_t A, C; make_masks(s, p, lane, A, C);
        NonEmptyRowSink sink; sink.begin(list, lane);
        bk_gen_all<NonEmptyRowSink, true>(A, C, sel4(s.used, p), sink);
        __syncwarp();
        const uint32_t ne = sink.count();                        // non-empty rows (uniform)
        const uint32_t per = (ne + 31u) >> 5;
        const uint32_t lo = min((uint32_t)lane * per, ne), hi = min(lo + per, ne);
        uint32_t cnt = 0;
        for (uint32_t e = lo; e < hi; e++) cnt += __popc(list[e] & BK_ROWMASK);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(BK_FULL, incl, d); if (lane >= d) incl += v; }
        const uint32_t total = __reduce_add_sync(BK_FULL, cnt);
        // CSR mode: state i owns ids[offsets[i] .. offsets[i+1]);  fixed mode (offsets == NULL): state i owns
        // ids[i*cap .. (i+1)*cap) and a longer list is truncated (counts_out still reports the full length).
        const size_t base = offsets ? (size_t)offsets[i] : (size_t)i * cap;
        const uint32_t limit = offsets ? total : (total < cap ? total : cap);
        if (counts_out && lane == 0) counts_out[i] = total;
        uint32_t pos = incl - cnt;                               // this lane's first list position
        uint16_t* out = ids + base;
        for (uint32_t e = lo; e < hi; e++) {
            const uint32_t w = list[e];
            const uint32_t id0 = (w >> 20) * 20u;
            uint32_t bits = w & BK_ROWMASK;
            do {                                                 // columns ascending = mirrored bits from the top
                const uint32_t j = 31u - (uint32_t)__clz(bits); bits ^= 1u << j;
                if (pos < limit) out[pos] = (uint16_t)(id0 + 19u - j);
                pos++;
            } while (bits);
        }
        __syncwarp();
    }
}

// The 1 820 row words of a state are staged in shared memory and leave as 455 full 16-byte stores per warp (7 280 B,
// 16-byte aligned per state); writing each orientation's 20 words straight from the sink made 91 partial-sector stores.
#define BITMAP_SMEM_PER_WARP (EMIT_ROW_WORDS * 4)
__global__ void __launch_bounds__(BK_THREADS)
k_movegen_bitmap(const bk_state* __restrict__ states, const uint8_t* __restrict__ players, int n,
                 uint32_t* __restrict__ bitmap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw + (size_t)wib * BITMAP_SMEM_PER_WARP);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint64_t word = warp < n ? fetch_state_word(states + warp, lane) : 0ull;
    for (int i = warp; i < n; i += nwarps) {
        const uint64_t next = (i + nwarps < n) ? fetch_state_word(states + i + nwarps, lane) : 0ull;   // prefetch
        WarpState s; unpack_state(word, lane, s);
        word = next;
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        uint32_t A, C; make_masks(s, p, lane, A, C);
#pragma unroll
        for (int k = 0; k < EMIT_WORDS_PER_LANE; k++) rows[k * 32 + lane] = 0;   // rows of spent pieces stay zero
        __syncwarp();
        SmemRowSink sink; sink.rows = rows; sink.lane = lane;
        bk_gen_all(A, C, sel4(s.used, p), sink);
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(bitmap + (size_t)i * BK_BITMAP_WORDS);
        const uint4* src = reinterpret_cast<const uint4*>(rows);
        for (int k = lane; k < BK_BITMAP_WORDS / 4; k += 32) __stcs(dst + k, src[k]);   // written once, never re-read here
        __syncwarp();
    }
}

// Per-piece placement counts P_i (the movegen-bound input of engine/mobility_metrics.py:38-131) and the largest
// number of legal moves sharing one anchor cell (its anchorTop1Share numerator).  Anchor counts are kept
// bit-sliced: plane k holds bit k of the count of every cell of this lane's row.
struct PieceCountSink {
    static constexpr bool kRows = false;            // counts per piece and per cell do not care about mirroring
    uint32_t acc = 0;
    uint32_t plane[7] = {0, 0, 0, 0, 0, 0, 0};      // <= 91 moves per anchor cell
    uint16_t* out; int lane;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) {
        acc += __popc(L);
        uint32_t carry = L;
#pragma unroll
        for (int k = 0; k < 7; k++) { const uint32_t t = plane[k] & carry; plane[k] ^= carry; carry = t; }
    }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {
        const uint32_t t = __reduce_add_sync(BK_FULL, acc);
        if (lane == 0) out[PID - 1] = (uint16_t)t;
        acc = 0;
    }
    __device__ __forceinline__ uint32_t anchor_max() const {
        uint32_t best = 0;
        for (int c = 0; c < 20; c++) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 7; k++) v |= ((plane[k] >> c) & 1u) << k;
            best = max(best, v);
        }
        return __reduce_max_sync(BK_FULL, best);
    }
};

__global__ void __launch_bounds__(BK_THREADS)
k_piece_counts(const bk_state* __restrict__ states, const uint8_t* __restrict__ players, int n,
               uint16_t* __restrict__ piece_counts, uint16_t* __restrict__ anchor_max) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        uint32_t A, C; make_masks(s, p, lane, A, C);
        uint16_t* out = piece_counts + (size_t)i * BK_PIECES;
        if (lane < BK_PIECES) out[lane] = 0;            // spent pieces are never visited
        __syncwarp();
        PieceCountSink sink; sink.out = out; sink.lane = lane;
        bk_gen_all<PieceCountSink, true>(A, C, sel4(s.used, p), sink);
        const uint32_t amax = sink.anchor_max();
        if (anchor_max && lane == 0) anchor_max[i] = (uint16_t)amax;
    }
}

__device__ __forceinline__ uint32_t alive_mask(const WarpState& s, int lane) {
    uint32_t mask = 0;
#pragma unroll 1
    for (uint32_t p = 0; p < 4; p++) {
        uint32_t A, C; make_masks(s, p, lane, A, C);
        // no corner cell at all -> no move (cheap early out, warp-uniform)
        if (!__any_sync(BK_FULL, C != 0)) continue;
        AnySink sink; bk_gen_all<AnySink, true>(A, C, sel4(s.used, p), sink);
        if (__any_sync(BK_FULL, sink.acc != 0)) mask |= 1u << p;
    }
    return mask;
}

__global__ void __launch_bounds__(BK_THREADS)
k_has_moves(const bk_state* __restrict__ states, int n, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        const uint32_t m = alive_mask(s, lane);
        if (lane == 0) out[i] = (uint8_t)m;
    }
}

// One move for `p`: piece unspent, orientation/anchor in range, bit set in the legal-anchor row.
__device__ __forceinline__ bool move_is_legal(const WarpState& s, uint32_t p, uint32_t id, int lane,
                                              int& o, int& row, int& col) {
    if (id >= (uint32_t)(BK_ORIENTS * BK_CELLS)) return false;
    o = id / 400; const int cell = id - o * 400; row = cell / 20; col = cell - row * 20;
    if (sel4(s.used, p) & (1u << (c_orient_piece[o] - 1))) return false;
    uint32_t A, C; make_masks(s, p, lane, A, C);
    const uint32_t L = bk_gen_one(A, C, o);
    return (__shfl_sync(BK_FULL, L, row) >> col) & 1u;
}

__global__ void __launch_bounds__(BK_THREADS)
k_is_move_legal(const bk_state* __restrict__ states, const uint8_t* __restrict__ players,
                const uint16_t* __restrict__ ids, int n, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        int o, row, col;
        const bool ok = move_is_legal(s, p, ids[i], lane, o, row, col);
        if (lane == 0) out[i] = ok ? 1 : 0;
    }
}

// Board.can_place_piece for an arbitrary cell set (engine/board.py:136-220): every cell empty and not
// edge-adjacent to an own cell (subset of A), and some cell on a corner cell / the start corner (hits C).
__global__ void __launch_bounds__(BK_THREADS)
k_can_place_mask(const bk_state* __restrict__ states, const uint8_t* __restrict__ players,
                 const uint64_t* __restrict__ masks, int n, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint64_t word = warp < n ? fetch_state_word(states + warp, lane) : 0ull;
    for (int i = warp; i < n; i += nwarps) {
        const uint64_t next = (i + nwarps < n) ? fetch_state_word(states + i + nwarps, lane) : 0ull;   // prefetch
        WarpState s; unpack_state(word, lane, s);
        word = next;
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        uint32_t A, C; make_masks(s, p, lane, A, C);
        const uint64_t word = lane < BK_WORDS ? masks[(size_t)i * BK_WORDS + lane] : 0ull;
        const int bitpos = lane * 20, wi = bitpos >> 6, off = bitpos & 63;
        const uint64_t lo = __shfl_sync(BK_FULL, word, wi & 31), hi = __shfl_sync(BK_FULL, word, (wi + 1) & 31);
        uint64_t v = lo >> off;
        if (off) v |= hi << (64 - off);
        const uint32_t m = lane < 20 ? ((uint32_t)v & BK_ROWMASK) : 0u;
        const bool inside = !__any_sync(BK_FULL, (m & ~A) != 0);
        const bool touches = __any_sync(BK_FULL, (m & C) != 0);
        if (lane == 0) out[i] = (inside && touches) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(BK_THREADS)
k_step(bk_state* __restrict__ states, const uint8_t* __restrict__ players, const uint16_t* __restrict__ ids,
       int n, uint8_t* __restrict__ okv, uint8_t* __restrict__ alive) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        const uint32_t id = ids[i];
        bool ok = true;
        if (id == BK_PASS) {
            s.to_move = (s.to_move + 1) & 3u;
        } else {
            const uint32_t p = players ? (players[i] & 3u) : s.to_move;
            int o, row, col;
            ok = move_is_legal(s, p, id, lane, o, row, col);
            if (ok) place_move(s, p, o, row, col, lane);
        }
        if (ok) store_state(states + i, lane, s);
        if (okv && lane == 0) okv[i] = ok ? 1 : 0;
        if (alive) {
            const uint32_t m = alive_mask(s, lane);
            if (lane == 0) alive[i] = (uint8_t)m;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: scores
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BK_THREADS)
k_scores(const bk_state* __restrict__ states, int n, int32_t* __restrict__ gs, int32_t* __restrict__ bs) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        int32_t g[4], b[4]; scores_of(s, lane, g, b);
        if (lane < 4) {
            const int32_t gv = lane == 0 ? g[0] : lane == 1 ? g[1] : lane == 2 ? g[2] : g[3];
            const int32_t bv = lane == 0 ? b[0] : lane == 1 ? b[1] : lane == 2 ? b[2] : b[3];
            if (gs) gs[(size_t)i * 4 + lane] = gv;
            if (bs) bs[(size_t)i * 4 + lane] = bv;
        }
    }
}

#include "heuristic.cuh"

// ---------------------------------------------------------------------------------------------
// K2: playouts.  One warp plays one whole game in registers.  POLICY 0 = uniform (agents/random_agent.py),
// POLICY 1 = HeuristicAgent softmax (agents/heuristic_agent.py, heuristic.cuh).
// ---------------------------------------------------------------------------------------------
#ifndef BK_PLAYOUT_MIN_BLOCKS
#define BK_PLAYOUT_MIN_BLOCKS 3
#endif
#ifndef BK_PLAYOUT_THREADS
#define BK_PLAYOUT_THREADS 256
#endif
#ifndef BK_PLAYOUT_MIRROR
#define BK_PLAYOUT_MIRROR true
#endif
#ifndef BK_PLAYOUT_HEUR_MIN_BLOCKS
#define BK_PLAYOUT_HEUR_MIN_BLOCKS 3
#endif
// POLICY 2 = one policy per seat and game (arena match-ups): seat_policies[i] packs the policy of player p in byte p.
template <int POLICY, bool TRACE = true>
__global__ void __launch_bounds__(BK_PLAYOUT_THREADS, POLICY == 0 ? BK_PLAYOUT_MIN_BLOCKS : BK_PLAYOUT_HEUR_MIN_BLOCKS)
k_playout(unsigned int* __restrict__ ticket, const uint32_t* __restrict__ seat_policies, const bk_state* __restrict__ start, int n, uint64_t seed, uint64_t game_offset, int first_turn, int max_turns,
                  int32_t* __restrict__ gs, int32_t* __restrict__ bs, uint16_t* __restrict__ trace,
                  uint16_t* __restrict__ choice, bk_state* __restrict__ finals) {
    __shared__ __align__(16) uint32_t s_slots[BK_PLAYOUT_THREADS / 32][BK_GROUP_SLOTS];
    extern __shared__ __align__(16) unsigned char heur_smem[];
    const uint32_t hs = (uint32_t)(threadIdx.x >> 5) * HEUR_SMEM_PER_WARP;     // this warp's heuristic scratch
    const int lane = threadIdx.x & 31;
    // Games differ in length (50..76 turns), so a fixed game->warp assignment leaves a tail of idle warps; every
    // warp instead takes the next game from a per-launch ticket counter.  Results do not depend on the assignment:
    // a game is a function of (seed, global game index) only.
    for (;;) {
        uint32_t i0 = 0;
        if (lane == 0) i0 = atomicAdd(ticket, 1u);
        const int i = (int)uni(i0);                        // warp-uniform game index (lanes != 0 contribute 0)
        if (i >= n) break;
        WarpState s;
        if (start) load_state(start + i, lane, s); else init_empty(s);
        const uint64_t g = game_offset + (uint64_t)i;
        const uint32_t seats = (POLICY == 2) ? uni(seat_policies[i]) : 0u;
        PhiloxBlock rnd; rnd.w[0] = rnd.w[1] = rnd.w[2] = rnd.w[3] = 0;
        uint16_t* tr = (TRACE && trace) ? trace + (size_t)i * BK_MAX_TURNS : nullptr;
        uint16_t* ch = (TRACE && choice) ? choice + (size_t)i * BK_MAX_TURNS : nullptr;
        for (int k = first_turn + lane; k < BK_MAX_TURNS; k += 32) { if (tr) tr[k] = BK_PASS; if (ch) ch[k] = BK_PASS; }
        __syncwarp();
        uint32_t dead = 0;
        // The reference loop stops the moment nobody can move (engine/game.py:182-214 runs after every
        // move and pass).  Here a player's death is discovered lazily on its own turn, which is the same
        // game as long as it runs to the end.  When max_turns cuts it short, up to four extra "probe" turns
        // (no placement) decide whether the game had in fact ended with the last placement, so that the
        // returned side-to-move equals the reference's in both cases.
        uint32_t resume_to_move = s.to_move;   // side to move right after the last placement
        uint32_t cut_to_move = s.to_move;      // side to move when max_turns was reached
        bool still_alive = false;
        for (uint32_t t = (uint32_t)first_turn; dead != 0xFu; t++) {
            const bool probing = t >= (uint32_t)max_turns;
            if (t == (uint32_t)max_turns) cut_to_move = s.to_move;
            if ((t & 3u) == 0u || t == (uint32_t)first_turn) rnd = philox_block(seed, g, t >> 2, 0);
            const uint32_t p = s.to_move;
            bool moved = false;
            if (!((dead >> p) & 1u)) {
                if (POLICY == 0 || (POLICY == 2 && ((seats >> (8u * p)) & 0xFFu) == BK_POLICY_UNIFORM)) {
                    uint32_t A, C; make_masks(s, p, lane, A, C);
                    GroupSink sink; sink.begin(s_slots[threadIdx.x >> 5], lane);
                    bk_gen_all<GroupSink, BK_PLAYOUT_MIRROR>(A, C, sel4(s.used, p), sink);
                    const Totals tot = reduce_totals(sink, lane);
                    if (tot.n == 0) {
                        dead |= 1u << p;      // a player without moves never regains any ("dead stay dead")
                    } else if (probing) {
                        still_alive = true;
                        break;
                    } else {
                        const uint32_t idx = __umulhi(philox_word(rnd, t), tot.n);
                        int o, row, col;
                        const uint32_t id = select_move(tot, idx, A, C, lane, o, row, col);
                        place_move(s, p, o, row, col, lane);
                        resume_to_move = s.to_move;
                        if (TRACE && lane == 0) { if (tr) tr[t] = (uint16_t)id; if (ch) ch[t] = (uint16_t)idx; }
                        moved = true;
                    }
                } else {
                    uint32_t id = 0, idx = 0;
                    const uint32_t nmoves = heuristic_pick(s, p, philox_word(rnd, t), hs, lane, ch != nullptr, id, idx);
                    if (nmoves == 0) {
                        dead |= 1u << p;
                    } else if (probing) {
                        still_alive = true;
                        break;
                    } else {
                        const int o = (int)(id / 400u), row = (int)((id % 400u) / 20u), col = (int)(id % 20u);
                        place_move(s, p, o, row, col, lane);
                        resume_to_move = s.to_move;
                        if (lane == 0) { if (tr) tr[t] = (uint16_t)id; if (ch) ch[t] = (uint16_t)idx; }
                        moved = true;
                    }
                }
            }
            if (!moved) s.to_move = (p + 1) & 3u;   // pass: engine/board.py:557-560 via arena_runner.py:662-666
        }
        s.to_move = still_alive ? cut_to_move : resume_to_move;
        int32_t gsc[4], bsc[4]; scores_of(s, lane, gsc, bsc);
        if (lane < 4) {
            const int32_t gv = lane == 0 ? gsc[0] : lane == 1 ? gsc[1] : lane == 2 ? gsc[2] : gsc[3];
            const int32_t bv = lane == 0 ? bsc[0] : lane == 1 ? bsc[1] : lane == 2 ? bsc[2] : bsc[3];
            if (gs) gs[(size_t)i * 4 + lane] = gv;
            if (bs) bs[(size_t)i * 4 + lane] = bv;
        }
        if (finals) store_state(finals + i, lane, s);
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of counts (three short kernels; n is at most a few million)
// ---------------------------------------------------------------------------------------------
#define SCAN_BLOCK 1024
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& block_total) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(BK_FULL, incl, d); if (lane >= d) incl += u; }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t s = warp_sums[lane], si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(BK_FULL, si, d); if (lane >= d) si += u; }
        warp_sums[lane] = si - s;
        if (lane == 31) block_total = si;
    }
    __syncthreads();
    const uint32_t res = warp_sums[w] + incl - v;
    return res;
}
// The first element of every 1024-block has within-block prefix 0, so its slot carries the block total
// between the passes: no scratch buffer, no allocation.
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_partial(const uint32_t* __restrict__ in, int n, uint32_t* __restrict__ out) {
    __shared__ uint32_t total;
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    const uint32_t ex = block_exclusive_scan(v, total);
    __syncthreads();
    if (i < n) out[i] = threadIdx.x == 0 ? total : ex;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_blocks(uint32_t* __restrict__ out, int nb, uint32_t* __restrict__ grand) {
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? out[(size_t)i * SCAN_BLOCK] : 0u;
        const uint32_t ex = block_exclusive_scan(v, total);
        if (i < nb) out[(size_t)i * SCAN_BLOCK] = ex + carry;
        __syncthreads();
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(uint32_t* __restrict__ out, int n) {
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n && threadIdx.x != 0) out[i] += out[(size_t)blockIdx.x * SCAN_BLOCK];
}

// ---------------------------------------------------------------------------------------------
// K4: Zobrist keys + open-addressing transposition table
// ---------------------------------------------------------------------------------------------
// mcts/zobrist.py:70-99: XOR over all 400 cells of table[row][col][value 0..4], the side to move and
// every spent piece of every player.
__device__ __forceinline__ uint64_t zobrist_of(const WarpState& s, const uint64_t* __restrict__ table, int lane) {
    uint64_t h = 0;
    if (lane < 20) {
        const uint64_t* rowt = table + (size_t)lane * 100;
#pragma unroll 4
        for (int c = 0; c < 20; c++) {
            const uint32_t v = ((s.P[0] >> c) & 1u) * 1u + ((s.P[1] >> c) & 1u) * 2u + ((s.P[2] >> c) & 1u) * 3u +
                               ((s.P[3] >> c) & 1u) * 4u;
            h ^= __ldg(rowt + c * 5 + v);
        }
    }
    if (lane < 21) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if ((s.used[q] >> lane) & 1u) h ^= __ldg(table + 2004 + q * 21 + lane);
    }
    if (lane == 31) h ^= __ldg(table + 2000 + s.to_move);
    uint32_t lo = __reduce_xor_sync(BK_FULL, (uint32_t)h), hi = __reduce_xor_sync(BK_FULL, (uint32_t)(h >> 32));
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(BK_THREADS)
k_zobrist(const bk_state* __restrict__ states, int n, const uint64_t* __restrict__ table, uint64_t* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        const uint64_t h = zobrist_of(s, table, lane);
        if (lane == 0) keys[i] = h;
    }
}

struct bk_tt {
    unsigned long long* keys;      // 0 = empty slot
    double* rewards;
    uint32_t* stamps;              // insertion number of the slot's entry (the dict order of the reference's table)
    uint32_t log2_slots;
    unsigned long long* occupied;  // device counters: [0] live entries, [1] entries ever inserted (next stamp),
                                   //                  [2] hits and [3] probes of the searches (hit_count / access_count)
    unsigned long long evicted;    // entries removed by bk_tt_evict_oldest: live stamps are [evicted, inserted)
    int device;
};
#define TT_MAX_PROBE 128
__device__ __forceinline__ uint64_t tt_norm(uint64_t k) { return k ? k : 0x9E3779B97F4A7C15ull; }
__device__ __forceinline__ uint32_t tt_home(uint64_t k, uint32_t log2_slots) {
    return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> (64 - log2_slots));
}
__device__ __forceinline__ bool tt_lookup(const unsigned long long* keys, const double* rewards, uint32_t log2_slots,
                                          uint64_t key, double& reward) {
    key = tt_norm(key);
    const uint32_t mask = (1u << log2_slots) - 1u;
    uint32_t slot = tt_home(key, log2_slots);
    for (int probe = 0; probe < TT_MAX_PROBE; probe++, slot = (slot + 1) & mask) {
        const unsigned long long k = keys[slot];
        if (k == key) { reward = rewards[slot]; return true; }
        if (k == 0ull) return false;
    }
    return false;
}
// stamp_in = 0xFFFFFFFF: a new entry takes the next insertion number; otherwise the given one (rebuild after eviction)
__device__ __forceinline__ bool tt_insert(unsigned long long* keys, double* rewards, uint32_t* stamps, uint32_t log2_slots,
                                          unsigned long long* occupied, uint64_t key, double reward,
                                          uint32_t stamp_in = 0xFFFFFFFFu) {
    key = tt_norm(key);
    const uint32_t mask = (1u << log2_slots) - 1u;
    uint32_t slot = tt_home(key, log2_slots);
    for (int probe = 0; probe < TT_MAX_PROBE; probe++, slot = (slot + 1) & mask) {
        unsigned long long k = keys[slot];
        if (k == 0ull) {
            k = atomicCAS(&keys[slot], 0ull, (unsigned long long)key);
            if (k == 0ull) {
                rewards[slot] = reward;
                stamps[slot] = stamp_in != 0xFFFFFFFFu ? stamp_in : (uint32_t)atomicAdd(occupied + 1, 1ull);
                atomicAdd(occupied, 1ull);
                return true;
            }
        }
        if (k == key) { rewards[slot] = reward; return true; }   // dict semantics: last put wins, the entry keeps its place
    }
    return false;   // neighbourhood full (the table holds 2x max_size slots; callers evict at max_size, see bk_tt_evict_oldest)
}
__global__ void k_tt_get(const unsigned long long* keys, const double* rewards, uint32_t log2_slots,
                         const uint64_t* q, int n, double* out, uint8_t* found) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = 0.0;
    const bool hit = tt_lookup(keys, rewards, log2_slots, q[i], r);
    found[i] = hit ? 1 : 0;
    if (hit) out[i] = r;
}
__global__ void k_tt_put(unsigned long long* keys, double* rewards, uint32_t* stamps, uint32_t log2_slots, unsigned long long* occupied,
                         const uint64_t* q, const double* r, int n) {
    if (gridDim.x == 1) {       // small batch: one thread, in order -- the puts enter the table in the order given
        if (threadIdx.x == 0)
            for (int i = 0; i < n; i++) tt_insert(keys, rewards, stamps, log2_slots, occupied, q[i], r[i]);
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // large batch: its entries count as inserted together
    if (i < n) tt_insert(keys, rewards, stamps, log2_slots, occupied, q[i], r[i]);
}
// TranspositionTable.put's eviction (mcts/zobrist.py:196-200): drop the entries with the lowest insertion numbers.
// Survivors are copied out, the table is wiped and they are inserted again with their insertion numbers.
__global__ void k_tt_collect(const unsigned long long* keys, const double* rewards, const uint32_t* stamps, uint32_t slots,
                             uint32_t first_kept, unsigned long long* tmp_keys, double* tmp_rewards, uint32_t* tmp_stamps,
                             unsigned int* n_kept) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots || keys[i] == 0ull) return;
    if ((int32_t)(stamps[i] - first_kept) < 0) return;            // older than the cut (wrap-safe)
    const unsigned int k = atomicAdd(n_kept, 1u);
    tmp_keys[k] = keys[i]; tmp_rewards[k] = rewards[i]; tmp_stamps[k] = stamps[i];
}
__global__ void k_tt_reinsert(unsigned long long* keys, double* rewards, uint32_t* stamps, uint32_t log2_slots,
                              unsigned long long* occupied, const unsigned long long* tmp_keys, const double* tmp_rewards,
                              const uint32_t* tmp_stamps, const unsigned int* n_kept) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_kept) return;
    tt_insert(keys, rewards, stamps, log2_slots, occupied, tmp_keys[i], tmp_rewards[i], tmp_stamps[i]);
}

// ---------------------------------------------------------------------------------------------
// N1: HeuristicAgent move scores (agents/heuristic_agent.py:68-198), float64, evaluated BEFORE the placement:
//   score = 1.0*size + 2.0*corners - 1.5*edge + 0.5*centre
//   corners = sum over piece cells, over their in-board diagonal neighbours d: d empty and no orthogonal
//             neighbour of d owned by the player  ==  d in the allowed mask A (duplicates counted, :118-138)
//   edge    = #cells with min(r, c, 19-r, 19-c) <= 2, halved once move_count/100 >= 0.3 (:153-173)
//   centre  = 1 - sqrt((ar-9.5)^2 + (ac-9.5)^2) / sqrt(2*9.5^2), from the ANCHOR (:185-198)
// Explicit _rn intrinsics keep the compiler from contracting to FMA, so the doubles equal numpy's bit for bit.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double heuristic_score(const uint32_t* __restrict__ arows, uint32_t move_count, uint32_t id,
                                                  double w_size = 1.0, double w_corner = 2.0, double w_edge = -1.5, double w_centre = 0.5) {
    const int o = id / 400, cell = id - o * 400, ar = cell / 20, ac = cell - ar * 20;
    const int n = c_orient_ncells[o];
    int corners = 0, edge = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t rc = c_orient_cells[o * 5 + i];
        const int r = ar + (int)(rc >> 4), c = ac + (int)(rc & 15);
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int nr = r + ((d & 2) ? 1 : -1), nc = c + ((d & 1) ? 1 : -1);
            if (nr >= 0 && nr < 20 && nc >= 0 && nc < 20) corners += (arows[nr] >> nc) & 1u;
        }
        const int m = min(min(r, c), min(19 - r, 19 - c));
        edge += (m <= 2);
    }
    double score = __dadd_rn(0.0, __dmul_rn(w_size, (double)n));
    score = __dadd_rn(score, __dmul_rn(w_corner, (double)corners));
    const double progress = __ddiv_rn((double)move_count, 100.0);
    const double edge_score = (progress < 0.3) ? (double)edge : __dmul_rn((double)edge, 0.5);
    score = __dadd_rn(score, __dmul_rn(w_edge, edge_score));
    const double dr = __dsub_rn((double)ar, 9.5), dc = __dsub_rn((double)ac, 9.5);
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc)));
    const double maxd = __dsqrt_rn(__dmul_rn(2.0, __dmul_rn(9.5, 9.5)));
    const double centre = __dsub_rn(1.0, __ddiv_rn(dist, maxd));
    return __dadd_rn(score, __dmul_rn(w_centre, centre));
}

__global__ void __launch_bounds__(BK_THREADS)
k_heuristic_scores(const bk_state* __restrict__ states, const uint8_t* __restrict__ players,
                   const uint32_t* __restrict__ offsets, const uint16_t* __restrict__ ids, int n,
                   double* __restrict__ out, double w_size, double w_corner, double w_edge, double w_centre) {
    __shared__ uint32_t s_a[BK_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint64_t word = warp < n ? fetch_state_word(states + warp, lane) : 0ull;
    for (int i = warp; i < n; i += nwarps) {
        const uint64_t next = (i + nwarps < n) ? fetch_state_word(states + i + nwarps, lane) : 0ull;   // prefetch
        WarpState s; unpack_state(word, lane, s);
        word = next;
        const uint32_t p = players ? (players[i] & 3u) : s.to_move;
        uint32_t A, C; make_masks(s, p, lane, A, C);
        __syncwarp();
        s_a[wib][lane] = A;
        __syncwarp();
        const uint32_t lo = offsets[i], hi = offsets[i + 1];
        for (uint32_t k = lo + lane; k < hi; k += 32)
            out[k] = heuristic_score(s_a[wib], s.move_count, ids[k], w_size, w_corner, w_edge, w_centre);
    }
}

#include "snapshot.cuh"
#include "mcts.cuh"
#include "mcts_wide.cuh"

#define BK_HEUR_PLAYOUT_SMEM ((BK_PLAYOUT_THREADS / 32) * HEUR_SMEM_PER_WARP)
static cudaError_t bk_init_policies_impl() {
    cudaError_t e = cudaFuncSetAttribute(k_playout<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_HEUR_PLAYOUT_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_playout<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_HEUR_PLAYOUT_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_mcts<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_MCTS_MAX_WARPS * HEUR_SMEM_PER_WARP);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_mcts<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEUR_SMEM_PER_WARP);
    if (e != cudaSuccess) return e;
    k_init_heuristic_tables<<<(91 * 400 + 255) / 256, 256>>>();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaDeviceSynchronize();
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
#define LAUNCH_CHECK() CUDA_TRY(cudaGetLastError())

extern "C" int bk_movegen_count(const bk_state* states, const uint8_t* players, int n, uint32_t* counts, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && counts)), "bk_movegen_count: null buffer");
    if (n == 0) return BK_OK;
    k_movegen_count<<<grid_for(cx, n, BK_COUNT_CTAS), BK_THREADS, 0, as_stream(stream)>>>(states, players, n, counts);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_movegen_emit(const bk_state* states, const uint8_t* players, int n, const uint32_t* offsets,
                               uint16_t* move_ids, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && offsets && move_ids)), "bk_movegen_emit: null buffer");
    if (n == 0) return BK_OK;
    k_movegen_emit<<<grid_for(cx, n, BK_EMIT_CTAS), BK_THREADS, BK_WARPS * EMIT_SMEM_PER_WARP, as_stream(stream)>>>(states, players, n, offsets, 0u, nullptr, move_ids);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_movegen_fixed(const bk_state* states, const uint8_t* players, int n, uint32_t cap, uint32_t* counts,
                                uint16_t* move_ids, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && cap > 0 && (n == 0 || (states && counts && move_ids)), "bk_movegen_fixed: bad argument");
    if (n == 0) return BK_OK;
    k_movegen_emit<<<grid_for(cx, n, BK_EMIT_CTAS), BK_THREADS, BK_WARPS * EMIT_SMEM_PER_WARP, as_stream(stream)>>>(states, players, n, nullptr, cap, counts, move_ids);
    LAUNCH_CHECK();
    return BK_OK;
}

// One board, host buffers in and out: the three steps of a single get_legal_moves call in one entry point (the copies are
// asynchronous when the host buffers are page-locked; the call returns after the stream has drained).
extern "C" int bk_legal_moves_host(const bk_state* state_host, bk_state* state_dev, void* list_dev, void* list_host, int capacity,
                                   void* stream) {
    REQUIRE_INIT();
    REQUIRE(state_host && state_dev && list_dev && list_host && capacity > 0 && capacity <= 32768, "bk_legal_moves_host: bad argument");
    cudaStream_t st = as_stream(stream);
    const size_t list_bytes = (size_t)capacity * sizeof(uint16_t);
    CUDA_TRY(cudaMemcpyAsync(state_dev, state_host, sizeof(bk_state), cudaMemcpyHostToDevice, st));
    k_movegen_emit<<<1, 32, EMIT_SMEM_PER_WARP, st>>>(state_dev, nullptr, 1, nullptr, (uint32_t)capacity,
                                                      reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(list_dev) + list_bytes),
                                                      static_cast<uint16_t*>(list_dev));
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(list_host, list_dev, list_bytes + 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return BK_OK;
}

// Fixed-capacity rows -> CSR: state i's first (offsets[i+1] - offsets[i]) ids move from slots[i*cap ...] to ids[offsets[i] ...].
__global__ void __launch_bounds__(BK_THREADS)
k_compact_rows(const uint16_t* __restrict__ slots, uint32_t cap, const uint32_t* __restrict__ offsets, int n,
               uint16_t* __restrict__ ids) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        const uint32_t lo = offsets[i], cnt = offsets[i + 1] - lo;
        const uint16_t* src = slots + (size_t)i * cap;
        for (uint32_t k = lane; k < cnt; k += 32) ids[lo + k] = __ldcs(src + k);      // read once
    }
}

extern "C" int bk_compact_rows(const uint16_t* slots, uint32_t cap, const uint32_t* offsets, int n, uint16_t* move_ids, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && cap > 0 && (n == 0 || (slots && offsets && move_ids)), "bk_compact_rows: bad argument");
    if (n == 0) return BK_OK;
    k_compact_rows<<<grid_for(cx, n, 8), BK_THREADS, 0, as_stream(stream)>>>(slots, cap, offsets, n, move_ids);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_movegen_bitmap(const bk_state* states, const uint8_t* players, int n, uint32_t* bitmap, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && bitmap)), "bk_movegen_bitmap: null buffer");
    REQUIRE(reinterpret_cast<uintptr_t>(bitmap) % 16 == 0, "bk_movegen_bitmap: bitmap must be 16-byte aligned");
    if (n == 0) return BK_OK;
    k_movegen_bitmap<<<grid_for(cx, n, 3), BK_THREADS, BK_WARPS * BITMAP_SMEM_PER_WARP, as_stream(stream)>>>(states, players, n, bitmap);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_movegen_piece_counts(const bk_state* states, const uint8_t* players, int n, uint16_t* piece_counts,
                                       uint16_t* anchor_max, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && piece_counts)), "bk_movegen_piece_counts: null buffer");
    if (n == 0) return BK_OK;
    k_piece_counts<<<grid_for(cx, n, 3), BK_THREADS, 0, as_stream(stream)>>>(states, players, n, piece_counts, anchor_max);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_snapshot_features(const bk_state* states, const uint64_t* frontiers, const int32_t* turn_index, int n,
                                    int max_turns, double* features, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && features)), "bk_snapshot_features: null buffer");
    if (n == 0) return BK_OK;
    k_snapshot_features<<<grid_for(cx, n, 3), BK_THREADS, BK_WARPS * BK_SNAP_SMEM_PER_WARP, as_stream(stream)>>>(
        states, frontiers, turn_index, n, max_turns, features);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_replay_snapshots(const bk_state* start, const uint16_t* traces, int n, int n_turns,
                                   const int32_t* checkpoints_host, int n_checkpoints, int max_turns, double* features,
                                   int32_t* turn_index, bk_state* snap_states, uint64_t* snap_frontiers, void* stream) {
    REQUIRE_INIT();
    REQUIRE(n >= 0 && n_turns >= 0 && (n == 0 || (traces && features && turn_index)), "bk_replay_snapshots: null buffer");
    REQUIRE(n_checkpoints >= 1 && n_checkpoints <= BK_MAX_CHECKPOINTS && checkpoints_host,
            "bk_replay_snapshots: 1..32 checkpoint plies required");
    bk_checkpoints ck; ck.n = n_checkpoints;
    for (int k = 0; k < BK_MAX_CHECKPOINTS; k++) ck.ply[k] = k < n_checkpoints ? checkpoints_host[k] : -1;
    for (int k = 0; k < n_checkpoints; k++)
        for (int j = 0; j < k; j++) REQUIRE(ck.ply[j] != ck.ply[k], "bk_replay_snapshots: duplicate checkpoint ply");
    if (n == 0) return BK_OK;
    k_replay_snapshots<<<grid_for(cx, n, 3), BK_THREADS, BK_WARPS * BK_SNAP_SMEM_PER_WARP, as_stream(stream)>>>(
        start, traces, n, n_turns, ck, max_turns, features, turn_index, snap_states, snap_frontiers);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_has_moves(const bk_state* states, int n, uint8_t* out_mask, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && out_mask)), "bk_has_moves: null buffer");
    if (n == 0) return BK_OK;
    k_has_moves<<<grid_for(cx, n, 3), BK_THREADS, 0, as_stream(stream)>>>(states, n, out_mask);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_is_move_legal(const bk_state* states, const uint8_t* players, const uint16_t* move_ids, int n,
                                uint8_t* out, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && move_ids && out)), "bk_is_move_legal: null buffer");
    if (n == 0) return BK_OK;
    k_is_move_legal<<<grid_for(cx, n, 6), BK_THREADS, 0, as_stream(stream)>>>(states, players, move_ids, n, out);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_can_place_mask(const bk_state* states, const uint8_t* players, const uint64_t* cell_masks, int n,
                                 uint8_t* out, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && cell_masks && out)), "bk_can_place_mask: null buffer");
    if (n == 0) return BK_OK;
    k_can_place_mask<<<grid_for(cx, n, 6), BK_THREADS, 0, as_stream(stream)>>>(states, players, cell_masks, n, out);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_step(bk_state* states, const uint8_t* players, const uint16_t* move_ids, int n, uint8_t* ok,
                       uint8_t* alive, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && move_ids)), "bk_step: null buffer");
    if (n == 0) return BK_OK;
    k_step<<<grid_for(cx, n, 3), BK_THREADS, 0, as_stream(stream)>>>(states, players, move_ids, n, ok, alive);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_scores(const bk_state* states, int n, int32_t* game_scores, int32_t* board_scores, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || states), "bk_scores: null buffer");
    if (n == 0) return BK_OK;
    k_scores<<<grid_for(cx, n, 6), BK_THREADS, 0, as_stream(stream)>>>(states, n, game_scores, board_scores);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_playout(const bk_state* start_states, int n, uint64_t seed, uint64_t game_offset, int policy,
                          int first_turn, int max_turns, int32_t* game_scores, int32_t* board_scores, uint16_t* trace,
                          uint16_t* choice, bk_state* final_states, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0, "bk_playout: negative n");
    REQUIRE(max_turns >= 0 && max_turns <= BK_MAX_TURNS, "bk_playout: max_turns out of range");
    if (policy != BK_POLICY_UNIFORM && policy != BK_POLICY_HEURISTIC) return fail(BK_ERR_ARG, "%s", "bk_playout: unknown policy");
    if (n == 0) return BK_OK;
    if (max_turns == 0) max_turns = BK_MAX_TURNS;
    REQUIRE(first_turn >= 0 && first_turn <= max_turns, "bk_playout: first_turn out of range");
    const int pw = BK_PLAYOUT_THREADS / 32;
    unsigned int* ticket = nullptr;
    CUDA_TRY(fresh_ticket(cx, as_stream(stream), &ticket));
    long long pneed = ((long long)n + pw - 1) / pw, pcap = (long long)cx->sms * BK_PLAYOUT_MIN_BLOCKS;
    if (policy == BK_POLICY_UNIFORM && !trace && !choice) {     // scores only: no per-turn output code at all
        k_playout<0, false><<<(int)(pneed < pcap ? pneed : pcap), BK_PLAYOUT_THREADS, 0, as_stream(stream)>>>(
            ticket, nullptr, start_states, n, seed, game_offset, first_turn, max_turns, game_scores, board_scores, trace, choice, final_states);
    } else if (policy == BK_POLICY_UNIFORM) {
        k_playout<0><<<(int)(pneed < pcap ? pneed : pcap), BK_PLAYOUT_THREADS, 0, as_stream(stream)>>>(
            ticket, nullptr, start_states, n, seed, game_offset, first_turn, max_turns, game_scores, board_scores, trace, choice, final_states);
    } else {
        pcap = (long long)cx->sms * BK_PLAYOUT_HEUR_MIN_BLOCKS;
        k_playout<1><<<(int)(pneed < pcap ? pneed : pcap), BK_PLAYOUT_THREADS, BK_HEUR_PLAYOUT_SMEM, as_stream(stream)>>>(
            ticket, nullptr, start_states, n, seed, game_offset, first_turn, max_turns, game_scores, board_scores, trace, choice, final_states);
    }
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_playout_seats(const bk_state* start_states, int n, uint64_t seed, uint64_t game_offset,
                                const uint32_t* seat_policies, int first_turn, int max_turns, int32_t* game_scores,
                                int32_t* board_scores, uint16_t* trace, uint16_t* choice, bk_state* final_states,
                                void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || seat_policies), "bk_playout_seats: null seat_policies");
    REQUIRE(max_turns >= 0 && max_turns <= BK_MAX_TURNS, "bk_playout_seats: max_turns out of range");
    if (n == 0) return BK_OK;
    if (max_turns == 0) max_turns = BK_MAX_TURNS;
    REQUIRE(first_turn >= 0 && first_turn <= max_turns, "bk_playout_seats: first_turn out of range");
    const int pw = BK_PLAYOUT_THREADS / 32;
    unsigned int* ticket = nullptr;
    CUDA_TRY(fresh_ticket(cx, as_stream(stream), &ticket));
    long long pneed = ((long long)n + pw - 1) / pw, pcap = (long long)cx->sms * BK_PLAYOUT_HEUR_MIN_BLOCKS;
    k_playout<2><<<(int)(pneed < pcap ? pneed : pcap), BK_PLAYOUT_THREADS, BK_HEUR_PLAYOUT_SMEM, as_stream(stream)>>>(
        ticket, seat_policies, start_states, n, seed, game_offset, first_turn, max_turns, game_scores, board_scores, trace, choice,
        final_states);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_exclusive_scan_u32(const uint32_t* counts, int n, uint32_t* offsets, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && offsets && (n == 0 || counts), "bk_exclusive_scan_u32: null buffer");
    cudaStream_t st = as_stream(stream);
    if (n == 0) { CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(uint32_t), st)); return BK_OK; }
    const int nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    k_scan_partial<<<nb, SCAN_BLOCK, 0, st>>>(counts, n, offsets);
    k_scan_blocks<<<1, SCAN_BLOCK, 0, st>>>(offsets, nb, offsets + n);
    k_scan_add<<<nb, SCAN_BLOCK, 0, st>>>(offsets, n);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_zobrist_hash(const bk_state* states, int n, const uint64_t* table, uint64_t* keys, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && table && (n == 0 || (states && keys)), "bk_zobrist_hash: null buffer");
    if (n == 0) return BK_OK;
    k_zobrist<<<grid_for(cx, n, 8), BK_THREADS, 0, as_stream(stream)>>>(states, n, table, keys);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_tt_create(bk_tt** out, uint32_t log2_slots) {
    REQUIRE_INIT(); REQUIRE(out && log2_slots >= 4 && log2_slots <= 30, "bk_tt_create: log2_slots must be in [4,30]");
    bk_tt* t = new (std::nothrow) bk_tt();
    if (!t) return fail(BK_ERR_ARG, "%s", "bk_tt_create: out of host memory");
    const size_t slots = (size_t)1 << log2_slots;
    t->log2_slots = log2_slots; t->keys = nullptr; t->rewards = nullptr; t->occupied = nullptr; t->stamps = nullptr; t->evicted = 0;
    cudaGetDevice(&t->device);
    cudaError_t e = cudaMalloc(&t->keys, slots * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&t->rewards, slots * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&t->stamps, slots * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&t->occupied, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(t->keys, 0, slots * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(t->occupied, 0, 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        cudaFree(t->keys); cudaFree(t->rewards); cudaFree(t->stamps); cudaFree(t->occupied); delete t;
        return fail(BK_ERR_CUDA, "bk_tt_create: %s", cudaGetErrorString(e));
    }
    *out = t;
    return BK_OK;
}
#define REQUIRE_TT_DEVICE(tt) \
    do { int d__ = -1; cudaGetDevice(&d__); if ((tt) && (tt)->device != d__) return fail(BK_ERR_ARG, "%s", "transposition table belongs to another CUDA device"); } while (0)
extern "C" int bk_tt_clear(bk_tt* tt, void* stream) {
    REQUIRE_INIT(); REQUIRE(tt, "bk_tt_clear: null table"); REQUIRE_TT_DEVICE(tt);
    CUDA_TRY(cudaMemsetAsync(tt->keys, 0, ((size_t)1 << tt->log2_slots) * sizeof(unsigned long long), as_stream(stream)));
    CUDA_TRY(cudaMemsetAsync(tt->occupied, 0, 4 * sizeof(unsigned long long), as_stream(stream)));
    tt->evicted = 0;
    return BK_OK;
}
extern "C" int bk_tt_destroy(bk_tt* tt) {
    if (!tt) return BK_OK;
    cudaFree(tt->keys); cudaFree(tt->rewards); cudaFree(tt->stamps); cudaFree(tt->occupied);
    delete tt;
    return BK_OK;
}
extern "C" int bk_tt_get(bk_tt* tt, const uint64_t* keys, int n, double* rewards, uint8_t* found, void* stream) {
    REQUIRE_INIT(); REQUIRE(tt && n >= 0 && (n == 0 || (keys && rewards && found)), "bk_tt_get: null buffer"); REQUIRE_TT_DEVICE(tt);
    if (n == 0) return BK_OK;
    k_tt_get<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(tt->keys, tt->rewards, tt->log2_slots, keys, n, rewards, found);
    LAUNCH_CHECK();
    return BK_OK;
}
extern "C" int bk_tt_put(bk_tt* tt, const uint64_t* keys, const double* rewards, int n, void* stream) {
    REQUIRE_INIT(); REQUIRE(tt && n >= 0 && (n == 0 || (keys && rewards)), "bk_tt_put: null buffer"); REQUIRE_TT_DEVICE(tt);
    if (n == 0) return BK_OK;
    k_tt_put<<<n <= 1024 ? 1 : (n + 255) / 256, n <= 1024 ? 32 : 256, 0, as_stream(stream)>>>(tt->keys, tt->rewards, tt->stamps, tt->log2_slots, tt->occupied, keys, rewards, n);
    LAUNCH_CHECK();
    return BK_OK;
}
extern "C" int bk_tt_stats(bk_tt* tt, uint64_t* out_host, void* stream) {
    REQUIRE_INIT(); REQUIRE(tt && out_host, "bk_tt_stats: null buffer"); REQUIRE_TT_DEVICE(tt);
    unsigned long long c[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(c, tt->occupied, sizeof(c), cudaMemcpyDeviceToHost, as_stream(stream)));
    CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
    out_host[0] = (uint64_t)1 << tt->log2_slots;
    out_host[1] = (uint64_t)c[0];
    out_host[2] = (uint64_t)c[2];
    out_host[3] = (uint64_t)c[3];
    return BK_OK;
}
extern "C" int bk_tt_evict_oldest(bk_tt* tt, uint64_t n_evict, void* stream) {
    REQUIRE_INIT(); REQUIRE(tt, "bk_tt_evict_oldest: null table"); REQUIRE_TT_DEVICE(tt);
    cudaStream_t st = as_stream(stream);
    unsigned long long c[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(c, tt->occupied, sizeof(c), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (n_evict == 0 || c[0] == 0) return BK_OK;
    if (n_evict >= c[0]) {                       // everything goes; hit / probe counters stay
        CUDA_TRY(cudaMemsetAsync(tt->keys, 0, ((size_t)1 << tt->log2_slots) * sizeof(unsigned long long), st));
        CUDA_TRY(cudaMemsetAsync(tt->occupied, 0, sizeof(unsigned long long), st));
        tt->evicted = c[1];
        return BK_OK;
    }
    const size_t keep = (size_t)(c[0] - n_evict);
    const uint32_t slots = 1u << tt->log2_slots;
    unsigned long long* tk = nullptr; double* tr = nullptr; uint32_t* ts = nullptr; unsigned int* nk = nullptr;
    cudaError_t e = cudaMalloc(&tk, keep * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&tr, keep * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&ts, keep * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&nk, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemsetAsync(nk, 0, sizeof(unsigned int), st);
    if (e == cudaSuccess) {
        const uint32_t first_kept = (uint32_t)(tt->evicted + n_evict);
        k_tt_collect<<<(slots + 255) / 256, 256, 0, st>>>(tt->keys, tt->rewards, tt->stamps, slots, first_kept, tk, tr, ts, nk);
        e = cudaMemsetAsync(tt->keys, 0, (size_t)slots * sizeof(unsigned long long), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(tt->occupied, 0, sizeof(unsigned long long), st);
        if (e == cudaSuccess) {
            k_tt_reinsert<<<(unsigned int)((keep + 255) / 256), 256, 0, st>>>(tt->keys, tt->rewards, tt->stamps, tt->log2_slots, tt->occupied, tk, tr, ts, nk);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(tk); cudaFree(tr); cudaFree(ts); cudaFree(nk);
    if (e != cudaSuccess) return fail(BK_ERR_CUDA, "bk_tt_evict_oldest: %s", cudaGetErrorString(e));
    tt->evicted += n_evict;
    return BK_OK;
}

extern "C" int64_t bk_mcts_workspace_bytes(const bk_mcts_params* params) {
    if (!params || params->iterations_total <= 0) return 0;
    return (int64_t)mcts_layout(params->iterations_total, params->reward_mode >= 3, params->max_depth).total;
}

extern "C" int bk_mcts_search(const bk_state* roots, int n_roots, int n_trees, const bk_mcts_params* params, bk_tt* tt,
                              const uint64_t* zobrist_table, void* workspace, int64_t workspace_bytes, int max_root_moves,
                              int32_t* root_visits, int64_t* root_reward, uint32_t* root_counts, uint16_t* root_move_ids,
                              uint32_t* status, void* stream) {
    REQUIRE_INIT();
    REQUIRE(roots && params && workspace && n_roots > 0 && n_trees > 0 && max_root_moves >= 0, "bk_mcts_search: bad argument");
    REQUIRE(params->iterations >= 0 && params->iterations_total > 0 && params->iterations_total <= 60000,
            "bk_mcts_search: iterations_total must be in [1, 60000]");
    REQUIRE(params->rollout_policy == BK_POLICY_UNIFORM || params->rollout_policy == BK_POLICY_HEURISTIC,
            "bk_mcts_search: unknown rollout policy");
    REQUIRE(!tt || zobrist_table, "bk_mcts_search: a transposition table needs a Zobrist table");
    const bool learned = params->reward_mode >= 3;
    const MctsLayout L = mcts_layout(params->iterations_total, learned, params->max_depth);
    if ((int64_t)L.total * n_trees > workspace_bytes) return fail(BK_ERR_CAPACITY, "%s", "bk_mcts_search: workspace too small");
    MctsTT t; t.keys = nullptr; t.rewards = nullptr; t.occupied = nullptr; t.stamps = nullptr; t.log2_slots = 0; t.zobrist = zobrist_table;
    if (tt) { REQUIRE_TT_DEVICE(tt); t.keys = tt->keys; t.rewards = tt->rewards; t.occupied = tt->occupied; t.stamps = tt->stamps; t.log2_slots = tt->log2_slots; }
    const bool heur = params->rollout_policy == BK_POLICY_HEURISTIC;
    const int W = params->leaf_batch > 1 ? params->leaf_batch : 1;
    REQUIRE(W <= BK_MCTS_MAX_WARPS, "bk_mcts_search: leaf_batch must be in [0, 16]");
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    long long* rr = reinterpret_cast<long long*>(root_reward);
    cudaStream_t st = as_stream(stream);
    REQUIRE(params->reward_mode >= 0 && params->reward_mode <= 4, "bk_mcts_search: unknown reward_mode");
    unsigned int* ticket = nullptr;
    CUDA_TRY(fresh_ticket(cx, st, &ticket));
    if (learned) {
        REQUIRE(params->leaf_model != nullptr, "bk_mcts_search: reward_mode 3/4 needs leaf_model (34 weights + bias on the device)");
        REQUIRE(W == 1, "bk_mcts_search: the learned modes run one warp per tree (leaf_batch must be 1)");
        const int cap = cx->sms * (heur && params->reward_mode == 4 ? 12 : 16);
        const int grid = n_trees < cap ? n_trees : cap;
        if (heur && params->reward_mode == 4)
            k_mcts<1, false, true><<<grid, BK_MCTS_THREADS, HEUR_SMEM_PER_WARP, st>>>(ticket, roots, n_roots, n_trees, *params, t,
                ws, params->resume, max_root_moves, root_visits, rr, root_counts, root_move_ids, status);
        else
            k_mcts<0, false, true><<<grid, BK_MCTS_THREADS, 0, st>>>(ticket, roots, n_roots, n_trees, *params, t,
                ws, params->resume, max_root_moves, root_visits, rr, root_counts, root_move_ids, status);
    } else if (W == 1) {
        const int cap = cx->sms * (heur ? BK_MCTS_HEUR_CTAS_PER_SM : BK_MCTS_CTAS_PER_SM);
        const int grid = n_trees < cap ? n_trees : cap;
        if (heur)
            k_mcts<1, false><<<grid, BK_MCTS_THREADS, HEUR_SMEM_PER_WARP, st>>>(ticket, roots, n_roots, n_trees, *params, t,
                ws, params->resume, max_root_moves, root_visits, rr, root_counts, root_move_ids, status);
        else
            k_mcts<0, false><<<grid, BK_MCTS_THREADS, 0, st>>>(ticket, roots, n_roots, n_trees, *params, t,
                ws, params->resume, max_root_moves, root_visits, rr, root_counts, root_move_ids, status);
    } else {
        const int cap = cx->sms * ((heur ? 16 : 24) / W);                       // CTAs of W warps: 16-24 warps per SM
        const int grid = n_trees < cap ? n_trees : cap;
        if (heur)
            k_mcts<1, true><<<grid, 32 * W, W * HEUR_SMEM_PER_WARP, st>>>(ticket, roots, n_roots, n_trees, *params, t,
                ws, params->resume, max_root_moves, root_visits, rr, root_counts, root_move_ids, status);
        else
            k_mcts<0, true><<<grid, 32 * W, 0, st>>>(ticket, roots, n_roots, n_trees, *params, t,
                ws, params->resume, max_root_moves, root_visits, rr, root_counts, root_move_ids, status);
    }
    LAUNCH_CHECK();
    return BK_OK;
}

static cudaError_t bk_init_wide() {
    return cudaFuncSetAttribute(k_wide_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_WIDE_PLAN_SMEM);
}

extern "C" int64_t bk_mcts_wide_workspace_bytes(const bk_mcts_params* params) {
    if (!params || params->iterations_total <= 0) return 0;
    return (int64_t)wide_layout(params->iterations_total, params->round_width, params->max_depth).total;
}

extern "C" int bk_mcts_wide_search(const bk_state* roots, int n_roots, int n_trees, const bk_mcts_params* params,
                                   void* workspace, int64_t workspace_bytes, int max_root_moves, int32_t* root_visits,
                                   int64_t* root_reward, uint32_t* root_counts, uint16_t* root_move_ids, uint32_t* status,
                                   uint32_t* sims_done, void* stream) {
    REQUIRE_INIT();
    REQUIRE(params && workspace && n_trees > 0 && max_root_moves >= 0 && sims_done, "bk_mcts_wide_search: bad argument");
    REQUIRE(params->resume || (roots && n_roots > 0), "bk_mcts_wide_search: a fresh search needs roots");
    REQUIRE(params->iterations >= 0 && params->iterations_total > 0 && params->iterations_total <= 1000000,
            "bk_mcts_wide_search: iterations_total must be in [1, 1000000]");
    REQUIRE(params->round_width >= 1 && params->round_width <= BK_WIDE_MAX_WIDTH, "bk_mcts_wide_search: round_width must be in [1, 4096]");
    REQUIRE(params->rollout_policy == BK_POLICY_UNIFORM || params->rollout_policy == BK_POLICY_HEURISTIC,
            "bk_mcts_wide_search: unknown rollout policy");
    REQUIRE(params->reward_mode >= 0 && params->reward_mode <= 2, "bk_mcts_wide_search: reward_mode must be 0, 1 or 2 (integer rewards)");
    REQUIRE(params->max_rollout_moves >= 0 && params->max_rollout_moves <= 96, "bk_mcts_wide_search: max_rollout_moves must be in [0, 96]");
    const WideLayout L = wide_layout(params->iterations_total, params->round_width, params->max_depth);
    if ((int64_t)L.total * n_trees > workspace_bytes) return fail(BK_ERR_CAPACITY, "%s", "bk_mcts_wide_search: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    const bool heur = params->rollout_policy == BK_POLICY_HEURISTIC;
    // simulations finished before this call (a resumed search continues where the header says)
    std::vector<uint32_t> done((size_t)n_trees, 0u);
    uint32_t before = 0;
    if (params->resume) {
        k_wide_finish<<<n_trees, 32, 0, st>>>(n_trees, *params, ws, 0, nullptr, nullptr, nullptr, nullptr, nullptr, sims_done);
        CUDA_TRY(cudaMemcpyAsync(done.data(), sims_done, sizeof(uint32_t) * (size_t)n_trees, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        before = done[0];
        for (int t = 1; t < n_trees; t++) REQUIRE(done[t] == before, "bk_mcts_wide_search: trees of one workspace must be resumed together");
    }
    const long long target_ll = (long long)before + params->iterations;
    REQUIRE(target_ll <= params->iterations_total, "bk_mcts_wide_search: search exceeds iterations_total");
    const uint32_t target = (uint32_t)target_ll;
    const int per_sm = heur ? BK_MCTS_HEUR_CTAS_PER_SM : BK_MCTS_CTAS_PER_SM;
    int gx = (cx->sms * per_sm + n_trees - 1) / n_trees;
    if (gx > (int)L.width) gx = (int)L.width;
    if (gx < 1) gx = 1;
    const dim3 wgrid((unsigned)gx, (unsigned)n_trees);
    int fresh = params->resume ? 0 : 1;
    uint32_t reached = before;
    if (fresh) {
        const unsigned tb = (L.table_len + 255u) / 256u;
        k_wide_tables<<<dim3(tb < 64u ? tb : 64u, (unsigned)n_trees), 256, 0, st>>>(n_trees, *params, ws);
    }
    // enqueue the rounds this call needs, then look at the counters (a tree that ran out of workspace stops early)
    for (int attempt = 0; attempt < 64 && reached < target; attempt++) {
        const uint32_t left = target - reached;
        int rounds = (int)((left + L.width - 1) / L.width);
        for (int r = 0; r < rounds; r++) {
            k_wide_plan<<<n_trees, BK_WIDE_PLAN_THREADS, BK_WIDE_PLAN_SMEM, st>>>(roots, n_roots > 0 ? n_roots : 1, n_trees, *params, ws, fresh, target);
            fresh = 0;
            if (heur) k_wide_work<1><<<wgrid, BK_MCTS_THREADS, HEUR_SMEM_PER_WARP, st>>>(n_trees, *params, ws);
            else k_wide_work<0><<<wgrid, BK_MCTS_THREADS, 0, st>>>(n_trees, *params, ws);
        }
        k_wide_finish<<<n_trees, 32, 0, st>>>(n_trees, *params, ws, max_root_moves, root_visits,
                                              reinterpret_cast<long long*>(root_reward), root_counts, root_move_ids, status, sims_done);
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemcpyAsync(done.data(), sims_done, sizeof(uint32_t) * (size_t)n_trees, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        uint32_t lo = done[0];
        for (int t = 1; t < n_trees; t++) lo = done[t] < lo ? done[t] : lo;
        if (lo == reached) break;             // no progress: a tree ran out of workspace (status says so) or has no moves
        reached = lo;
    }
    if (params->iterations == 0 || fresh) {    // nothing ran: still build the root (fresh) and report
        if (fresh) k_wide_plan<<<n_trees, BK_WIDE_PLAN_THREADS, BK_WIDE_PLAN_SMEM, st>>>(roots, n_roots, n_trees, *params, ws, 1, 0u);
        k_wide_finish<<<n_trees, 32, 0, st>>>(n_trees, *params, ws, max_root_moves, root_visits,
                                              reinterpret_cast<long long*>(root_reward), root_counts, root_move_ids, status, sims_done);
        LAUNCH_CHECK();
    }
    return BK_OK;
}

extern "C" int bk_heuristic_scores(const bk_state* states, const uint8_t* players, const uint32_t* offsets,
                                   const uint16_t* move_ids, int n, double* scores, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && offsets && move_ids && scores)), "bk_heuristic_scores: null buffer");
    if (n == 0) return BK_OK;
    k_heuristic_scores<<<grid_for(cx, n, 6), BK_THREADS, 0, as_stream(stream)>>>(states, players, offsets, move_ids, n, scores, 1.0, 2.0, -1.5, 0.5);
    LAUNCH_CHECK();
    return BK_OK;
}

extern "C" int bk_heuristic_scores_weighted(const bk_state* states, const uint8_t* players, const uint32_t* offsets,
                                            const uint16_t* move_ids, int n, const double* weights_host, double* scores, void* stream) {
    REQUIRE_INIT(); REQUIRE(n >= 0 && (n == 0 || (states && offsets && move_ids && scores)), "bk_heuristic_scores_weighted: null buffer");
    REQUIRE(weights_host != nullptr, "bk_heuristic_scores_weighted: weights_host = {piece_size, corner_creation, edge_avoidance, center_preference}");
    if (n == 0) return BK_OK;
    k_heuristic_scores<<<grid_for(cx, n, 6), BK_THREADS, 0, as_stream(stream)>>>(states, players, offsets, move_ids, n, scores,
        weights_host[0], weights_host[1], weights_host[2], weights_host[3]);
    LAUNCH_CHECK();
    return BK_OK;
}

// ---------------------------------------------------------------------------------------------
// Root-parallel merge over peer memory (one process per GPU; windows mapped through CUDA IPC)
// ---------------------------------------------------------------------------------------------
// Window layout in 64-bit words: two slots of [visits[capacity], reward[capacity]] (merge e uses slot e & 1), then the two
// arrival counters.  A slot is cleared by its owner after it has been read; it is written again two merges later, and a
// peer can only be that far ahead after the owner has taken part in the merge in between, i.e. after this kernel ended.
#define BK_P2P_MAX_WORLD 16
struct P2PPeers { unsigned long long* w[BK_P2P_MAX_WORLD]; };

__global__ void __launch_bounds__(1024)
k_root_merge_p2p(const void* __restrict__ visits, int visits_i64, const int64_t* __restrict__ reward, int n_trees, int stride, int n, int capacity,
                 int rank, int world, P2PPeers peers, unsigned long long epoch, long long timeout_cycles,
                 int64_t* __restrict__ merged, int32_t* __restrict__ status) {
    __shared__ int s_ok;
    const unsigned slot = (unsigned)(epoch & 1ull);
    const unsigned long long target = (epoch / 2ull + 1ull) * (unsigned long long)world;
    const size_t slot_base = (size_t)slot * 2u * (size_t)capacity;
    // 1. this rank's sums over its trees, added into every rank's window (own included)
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
        const int which = i >= n, k = i - which * n;
        long long v = 0;
        for (int t = 0; t < n_trees; t++)
            v += which ? (long long)reward[(size_t)t * stride + k]
                       : (visits_i64 ? (long long)static_cast<const int64_t*>(visits)[(size_t)t * stride + k]
                                     : (long long)static_cast<const int32_t*>(visits)[(size_t)t * stride + k]);
        const size_t at = slot_base + (size_t)which * capacity + k;
        for (int p = 0; p < world; p++) {
            const int q = (rank + 1 + p) % world;                        // ranks start with different peers
            atomicAdd_system(peers.w[q] + at, (unsigned long long)v);
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. arrival: one increment per rank and window; wait for `world` arrivals at the own window
    if (threadIdx.x == 0) {
        for (int p = 0; p < world; p++) atomicAdd_system(peers.w[(rank + 1 + p) % world] + 4u * (size_t)capacity + slot, 1ull);
        volatile unsigned long long* mine = peers.w[rank] + 4u * (size_t)capacity + slot;
        const long long t0 = clock64();
        int ok = 1;
        while (*mine < target) {
            if (clock64() - t0 > timeout_cycles) { ok = 0; break; }
            __nanosleep(200);
        }
        s_ok = ok;
    }
    __syncthreads();
    __threadfence_system();
    // 3. read the merged sums and clear the slot for the merge after the next
    volatile unsigned long long* own = peers.w[rank];
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
        const int which = i >= n, k = i - which * n;
        const size_t at = slot_base + (size_t)which * capacity + k;
        merged[(size_t)which * n + k] = (int64_t)own[at];
        own[at] = 0ull;
    }
    __threadfence_system();
    if (threadIdx.x == 0) *status = s_ok ? 0 : 1;
}

extern "C" int bk_p2p_window_create(int capacity, void** window, unsigned char* handle_host) {
    REQUIRE_INIT(); REQUIRE(window && handle_host, "bk_p2p_window_create: null output");
    REQUIRE(capacity >= 1 && capacity <= 2048, "bk_p2p_window_create: capacity must be in [1, 2048]");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    const size_t bytes = (4u * (size_t)capacity + 2u) * sizeof(unsigned long long);
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p); return fail(BK_ERR_CUDA, "bk_p2p_window_create: %s", cudaGetErrorString(e)); }
    memcpy(handle_host, &h, sizeof(h));
    *window = p;
    return BK_OK;
}
extern "C" int bk_p2p_window_open(const unsigned char* handle_host, void** peer_window) {
    REQUIRE_INIT(); REQUIRE(handle_host && peer_window, "bk_p2p_window_open: null argument");
    cudaIpcMemHandle_t h; memcpy(&h, handle_host, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(peer_window, h, cudaIpcMemLazyEnablePeerAccess));
    return BK_OK;
}
extern "C" int bk_p2p_window_close(void* peer_window) {
    if (peer_window) CUDA_TRY(cudaIpcCloseMemHandle(peer_window));
    return BK_OK;
}
extern "C" int bk_p2p_window_destroy(void* window) {
    if (window) CUDA_TRY(cudaFree(window));
    return BK_OK;
}
extern "C" int bk_root_merge_p2p(const void* root_visits, int visits_are_int64, const int64_t* root_reward, int n_trees, int stride, int n, int capacity,
                                 int rank, int world, void* const* windows_host, uint64_t epoch, int timeout_ms, int64_t* merged,
                                 int32_t* status, void* stream) {
    REQUIRE_INIT();
    REQUIRE(root_visits && root_reward && windows_host && merged && status, "bk_root_merge_p2p: null buffer");
    REQUIRE(world >= 1 && world <= BK_P2P_MAX_WORLD && rank >= 0 && rank < world, "bk_root_merge_p2p: rank / world out of range (world <= 16)");
    REQUIRE(n_trees >= 1 && n >= 0 && n <= capacity && n <= stride && capacity <= 2048, "bk_root_merge_p2p: n, stride or capacity out of range");
    P2PPeers peers;
    for (int p = 0; p < BK_P2P_MAX_WORLD; p++) peers.w[p] = p < world ? reinterpret_cast<unsigned long long*>(windows_host[p]) : nullptr;
    for (int p = 0; p < world; p++) REQUIRE(peers.w[p] != nullptr, "bk_root_merge_p2p: null window");
    static int khz_of[BK_MAX_DEVICES];                  // cudaDevAttrClockRate is a slow query (about a millisecond): asked once per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (khz_of[dev] == 0) {
        int q = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&q, cudaDevAttrClockRate, dev));
        khz_of[dev] = q > 0 ? q : 2000000;
    }
    const int khz = khz_of[dev];
    const long long timeout_cycles = (long long)(timeout_ms > 0 ? timeout_ms : 20000) * (long long)(khz > 0 ? khz : 2000000);
    k_root_merge_p2p<<<1, 1024, 0, as_stream(stream)>>>(root_visits, visits_are_int64, root_reward, n_trees, stride, n, capacity, rank, world, peers,
                                                        (unsigned long long)epoch, timeout_cycles, merged, status);
    LAUNCH_CHECK();
    return BK_OK;
}
