// movegen_core.cuh -- row-sliced Blokus legality on one warp per game state (sm_100a).
//
// Replaces the per-candidate loop of the reference's generator
// (engine/move_generator.py:261-559 calling is_placement_legal_bitboard_fast :760-831, which in turn
// calls shift_mask_fast engine/bitboard.py:173-205 three times per candidate) by a dense bit-parallel
// evaluation: lane r (0..19) of the warp owns row r of every 20x20 mask as the low 20 bits of a
// register; lanes 20..31 hold zero rows.  A row shift is one SHFL, a column shift one SHF, and bits
// 20..31 / lanes 20..31 being zero makes "off the board" fall out as "not allowed" with no wrap masks.
//
//   own  = mover's cells, occ = all cells
//   A    = ~(occ | N(own)|S(own)|E(own)|W(own))          cells the mover may cover (rules 1+2)
//   C    = first move ? A & start_corner : A & (NE|NW|SE|SW)(own)   cells that satisfy rule 3
//   legal_o(row, col) = AND_i A[row+r_i][col+c_i]  &  OR_i C[row+r_i][col+c_i]
//
// The 91 orientation bodies are generated (orient_gen.cuh) with shared run primitives.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/blokus_b200.h"
#include "tables_gen.h"

#define BK_FULL 0xffffffffu
#define BK_ROWMASK 0xFFFFFu
#define BK_ALL_PIECES 0x1FFFFFu

__constant__ uint8_t c_orient_piece[BK_TBL_ORIENTS];
__constant__ uint8_t c_orient_ncells[BK_TBL_ORIENTS];
__constant__ uint8_t c_orient_cells[BK_TBL_ORIENTS * 5];
__constant__ uint8_t c_group_first[BK_TBL_GROUPS + 1];
__constant__ uint32_t c_orient_cells32[BK_TBL_ORIENTS];   // 5 x (row<<3|col), padded with repeats of cell 0
__constant__ uint32_t c_orient_rowpat[BK_TBL_ORIENTS];    // 5 rows x 5 column bits
__constant__ uint8_t c_piece_first[BK_TBL_PIECES + 2];

// ---- per-warp game state in registers ----------------------------------------------------------
struct WarpState {
    uint32_t P[4];     // this lane's row of each player's cells
    uint32_t used[4];  // warp-uniform
    uint32_t to_move, first_mask, move_count;  // warp-uniform
};

__device__ __forceinline__ uint32_t sel4(const uint32_t (&v)[4], uint32_t i) {
    uint32_t a = (i & 1) ? v[1] : v[0];
    uint32_t b = (i & 1) ? v[3] : v[2];
    return (i & 2) ? b : a;
}

// The scalars of a state are equal in all lanes; passing them through REDUX once tells the compiler so (the
// results live in uniform registers), and branches on them need no reconvergence bookkeeping.
__device__ __forceinline__ uint32_t uni(uint32_t x) { return __reduce_max_sync(BK_FULL, x); }
__device__ __forceinline__ void make_uniform(WarpState& s) {
#pragma unroll
    for (int q = 0; q < 4; q++) s.used[q] = uni(s.used[q]);
    s.to_move = uni(s.to_move); s.first_mask = uni(s.first_mask); s.move_count = uni(s.move_count);
}

__device__ __forceinline__ void init_empty(WarpState& s) {
#pragma unroll
    for (int q = 0; q < 4; q++) { s.P[q] = 0; s.used[q] = 0; }
    s.to_move = 0; s.first_mask = 0xF; s.move_count = 0;
}

// 256-byte packed state -> rows.  Lane i reads 64-bit word i (two coalesced 128-byte lines per warp).  The
// fetch is separate from the unpacking so that streaming kernels can issue the next state's load before they
// work on the current one (the count kernel was waiting on this load: long_scoreboard 1.6 per issue, r01d).
__device__ __forceinline__ uint64_t fetch_state_word(const bk_state* __restrict__ sp, int lane) {
    return __ldg(reinterpret_cast<const uint64_t*>(sp) + lane);
}
__device__ __forceinline__ void unpack_state(const uint64_t word, int lane, WarpState& s) {
    const int bitpos = lane * 20;
    const int wi = bitpos >> 6, off = bitpos & 63;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint64_t lo = __shfl_sync(BK_FULL, word, (q * 7 + wi) & 31);
        uint64_t hi = __shfl_sync(BK_FULL, word, (q * 7 + wi + 1) & 31);
        uint64_t v = lo >> off;
        if (off) v |= hi << (64 - off);
        s.P[q] = (lane < 20) ? ((uint32_t)v & BK_ROWMASK) : 0u;
    }
    uint64_t u01 = __shfl_sync(BK_FULL, word, 28), u23 = __shfl_sync(BK_FULL, word, 29);
    uint64_t meta = __shfl_sync(BK_FULL, word, 30);
    s.used[0] = (uint32_t)u01; s.used[1] = (uint32_t)(u01 >> 32);
    s.used[2] = (uint32_t)u23; s.used[3] = (uint32_t)(u23 >> 32);
    s.to_move = (uint32_t)meta & 3u;
    s.first_mask = (uint32_t)(meta >> 8) & 0xFu;
    s.move_count = (uint32_t)(meta >> 16) & 0xFFFFu;
    make_uniform(s);
}
__device__ __forceinline__ void load_state(const bk_state* __restrict__ sp, int lane, WarpState& s) {
    unpack_state(fetch_state_word(sp, lane), lane, s);
}

// rows -> 256-byte packed state.  Lane i assembles 64-bit word i.
__device__ __forceinline__ void store_state(bk_state* __restrict__ sp, int lane, const WarpState& s) {
    uint64_t word = 0;
    const int q = lane / 7, w = lane - q * 7;         // valid for lane < 28
    const int r0 = (64 * w) / 20;                      // first row overlapping this word
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int rr = r0 + j;
        uint32_t v0 = __shfl_sync(BK_FULL, s.P[0], rr & 31);
        uint32_t v1 = __shfl_sync(BK_FULL, s.P[1], rr & 31);
        uint32_t v2 = __shfl_sync(BK_FULL, s.P[2], rr & 31);
        uint32_t v3 = __shfl_sync(BK_FULL, s.P[3], rr & 31);
        uint32_t v = (q == 0) ? v0 : (q == 1) ? v1 : (q == 2) ? v2 : v3;
        const int d = 20 * rr - 64 * w;                // bit offset of row rr inside this word
        if (rr < 20 && d < 64) word |= (d >= 0) ? ((uint64_t)v << d) : ((uint64_t)v >> (-d));
    }
    if (lane == 28) word = (uint64_t)s.used[0] | ((uint64_t)s.used[1] << 32);
    if (lane == 29) word = (uint64_t)s.used[2] | ((uint64_t)s.used[3] << 32);
    if (lane == 30) word = (uint64_t)(s.to_move & 3u) | ((uint64_t)(s.first_mask & 0xFu) << 8) |
                           ((uint64_t)(s.move_count & 0xFFFFu) << 16);
    if (lane == 31) word = 0;
    reinterpret_cast<uint64_t*>(sp)[lane] = word;
}

// allowed (A) and corner (C) rows for `player` -- the rules of engine/board.py:136-220.
__device__ __forceinline__ void make_masks(const WarpState& s, uint32_t player, int lane,
                                           uint32_t& A, uint32_t& C) {
    const uint32_t own = sel4(s.P, player);
    const uint32_t occ = s.P[0] | s.P[1] | s.P[2] | s.P[3];
    const uint32_t h = ((own << 1) | (own >> 1)) & BK_ROWMASK;      // E/W neighbours
    uint32_t up_own = __shfl_up_sync(BK_FULL, own, 1), dn_own = __shfl_down_sync(BK_FULL, own, 1);
    uint32_t up_h = __shfl_up_sync(BK_FULL, h, 1), dn_h = __shfl_down_sync(BK_FULL, h, 1);
    if (lane == 0) { up_own = 0; up_h = 0; }
    if (lane >= 19) { dn_own = 0; dn_h = 0; }
    const uint32_t orth = h | up_own | dn_own;
    const uint32_t diag = up_h | dn_h;
    A = (lane < 20) ? (~(occ | orth) & BK_ROWMASK) : 0u;
    if ((s.first_mask >> player) & 1u) {
        // start corners RED(0,0) BLUE(0,19) YELLOW(19,19) GREEN(19,0): engine/board.py:56-61
        const int crow = (player >= 2) ? 19 : 0;
        const uint32_t cbit = (player == 1 || player == 2) ? (1u << 19) : 1u;
        C = (lane == crow) ? (A & cbit) : 0u;
    } else {
        C = A & diag;
    }
}

// Evaluate every orientation of every unspent piece; sink.put<O>(row_of_legal_anchors).
//
// MIRROR = true runs the generated code on column-mirrored rows (bit j = column 19-j), which turns every
// column shift into a LEFT shift; ptxas issues those as IMAD.SHL on the FMA pipe instead of SHF on the ALU pipe
// that LOP3 already loads (profiles/r01a_playout_ncu.md: ALU 69 %, FMA 10 %).  Garbage above bit 19 cannot
// survive: every orientation has a cell in column 0, so its AND contains an unshifted, clean factor.  Measured
// on B200: +8 % for the count kernel, +6 % for the playout kernel (14.6 -> 15.5 M playouts/s), so the counting
// sinks use it.  Sinks that store rows (kRows) get them mirrored back with one BREV.
template <class Sink, bool MIRROR = false>
__device__ __forceinline__ void bk_gen_all(const uint32_t A_in, const uint32_t C_in, const uint32_t used_in, Sink& sink) {
    // `used` is the same in every lane; routing it through REDUX puts it in a uniform register, so the 21 piece
    // tests become uniform branches (no BSSY/BSYNC reconvergence pairs around every piece block)
    const uint32_t used = __reduce_or_sync(BK_FULL, used_in);
    const uint32_t A = MIRROR ? (__brev(A_in) >> 12) : A_in, C = MIRROR ? (__brev(C_in) >> 12) : C_in;
#define BK_CS(x, c) (MIRROR ? ((x) << (c)) : ((x) >> (c)))
#define BK_OUT(L) ((MIRROR && Sink::kRows) ? (__brev(L) >> 12) : (L))
#include "orient_gen.cuh"
#undef BK_CS
#undef BK_OUT
}

// One orientation, table-driven (used after a move has been chosen, and by is_move_legal).  Short shapes are
// padded by repeating their first cell (AND / OR are idempotent), so the loop has no shape-dependent branch.
__device__ __forceinline__ uint32_t bk_gen_one(const uint32_t A, const uint32_t C, const int o) {
    const uint32_t cells = c_orient_cells32[o];
    uint32_t fit = BK_ROWMASK, touch = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint32_t rc = (cells >> (6 * i)) & 63u;
        fit &= __shfl_down_sync(BK_FULL, A, rc >> 3) >> (rc & 7u);
        touch |= __shfl_down_sync(BK_FULL, C, rc >> 3) >> (rc & 7u);
    }
    return fit & touch;
}

// ---- sinks ------------------------------------------------------------------------------------
// put<O, SLOT>(L): this lane's row of legal anchors of orientation O (SLOT = index of O in its count group);
// flush<G>(): end of count group G (emitted inside the piece's warp-uniform branch).
struct CountSink {                 // per-lane move count
    static constexpr bool kRows = false;
    uint32_t acc = 0;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) { acc += __popc(L); }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
};

struct AnySink {                   // does any move exist
    static constexpr bool kRows = false;
    uint32_t acc = 0;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) { acc |= L; }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
};

struct BitmapSink {                // legal-anchor bitmap rows straight to memory
    static constexpr bool kRows = true;
    uint32_t* out; int lane;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) {
        if (lane < 20) out[O * 20 + lane] = L;
    }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
};

// Per-orientation totals for choosing the idx-th move.  The counts of a group's <= 3 orientations sit in 10-bit
// fields of ONE register (a row holds <= 20 anchors, a warp-wide field sum <= 400 < 1024); when the group ends a
// single REDUX sums them over the warp (the result lands in a uniform register) and lane 0 parks it in the
// warp's shared-memory slot G.  Only one accumulator is live at a time, and the bookkeeping costs no ALU-pipe
// instruction (REDUX on the ADU, one predicated STS).
#define BK_GROUP_SLOTS 44            /* 41 groups, padded to a multiple of 4 words; the last slot is a dump */
struct GroupSink {
    static constexpr bool kRows = false;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0;   // the (up to three) groups of the piece being evaluated
    uint32_t* slots;                          // this warp's BK_GROUP_SLOTS words of shared memory, zeroed by begin()
    int lane;
    __device__ __forceinline__ void begin(uint32_t* warp_slots, int lane_) {
        slots = warp_slots; lane = lane_;
        __syncwarp();
        if (lane < BK_GROUP_SLOTS / 2) reinterpret_cast<uint2*>(slots)[lane] = make_uint2(0u, 0u);
        __syncwarp();
    }
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) {
        const uint32_t v = (uint32_t)__popc(L) << (10 * SLOT);
        if (GI == 0) acc0 += v; else if (GI == 1) acc1 += v; else acc2 += v;
    }
    // All REDUX of a piece are issued back to back (their latencies overlap) before the first result is stored.
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {
        const uint32_t t0 = __reduce_add_sync(BK_FULL, acc0);
        const uint32_t t1 = NG > 1 ? __reduce_add_sync(BK_FULL, acc1) : 0u;
        const uint32_t t2 = NG > 2 ? __reduce_add_sync(BK_FULL, acc2) : 0u;
        if (lane == 0) {
            slots[G0] = t0;
            if (NG > 1) slots[G0 + 1] = t1;
            if (NG > 2) slots[G0 + 2] = t2;
        }
        acc0 = 0; if (NG > 1) acc1 = 0; if (NG > 2) acc2 = 0;
    }
    __device__ __forceinline__ void finish() {}
};

// ---- choosing the idx-th legal move in ascending id order ----------------------------------------
struct Totals {
    uint32_t mine0, mine1;   // packed totals of groups 2*lane and 2*lane+1
    uint32_t incl;           // inclusive prefix over lanes of this lane's move count
    uint32_t n;              // total number of legal moves (uniform)
};

__device__ __forceinline__ uint32_t fields_sum(uint32_t m) { return (m & 1023u) + ((m >> 10) & 1023u) + ((m >> 20) & 1023u); }

__device__ __forceinline__ Totals reduce_totals(GroupSink& s, int lane) {
    Totals t;
    s.finish();
    __syncwarp();
    const uint2 m = (lane < BK_GROUP_SLOTS / 2) ? reinterpret_cast<const uint2*>(s.slots)[lane] : make_uint2(0u, 0u);
    t.mine0 = m.x; t.mine1 = m.y;
    const uint32_t mysum = fields_sum(m.x) + fields_sum(m.y);
    uint32_t incl = mysum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(BK_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    t.incl = incl;
    t.n = __reduce_add_sync(BK_FULL, mysum);   // uniform register: `n == 0` becomes a uniform branch
    return t;
}

// Locate move number idx (0-based, idx < t.n): returns the move id; A/C are the masks used for the count.
__device__ __forceinline__ uint32_t select_move(const Totals& t, uint32_t idx, uint32_t A, uint32_t C, int lane,
                                                int& o_out, int& row_out, int& col_out) {
    // lane holding the group pair that contains idx
    const uint32_t vote = __ballot_sync(BK_FULL, t.incl > idx);
    const int j = __ffs(vote) - 1;
    const uint32_t m0 = __shfl_sync(BK_FULL, t.mine0, j), m1 = __shfl_sync(BK_FULL, t.mine1, j);
    const uint32_t incl = __shfl_sync(BK_FULL, t.incl, j);
    const uint32_t s0 = fields_sum(m0), s1 = fields_sum(m1);
    uint32_t rem = idx - (incl - s0 - s1);
    int g = 2 * j; uint32_t m = m0;
    if (rem >= s0) { rem -= s0; g++; m = m1; }
    const uint32_t f0 = m & 1023u, f1 = (m >> 10) & 1023u;
    int o = c_group_first[g];
    if (rem >= f0) { rem -= f0; o++; if (rem >= f1) { rem -= f1; o++; } }
    // rows of that orientation
    const uint32_t L = bk_gen_one(A, C, o);
    const uint32_t c = __popc(L);
    uint32_t rincl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(BK_FULL, rincl, d);
        if (lane >= d) rincl += v;
    }
    const uint32_t rvote = __ballot_sync(BK_FULL, rincl > rem);
    const int row = __ffs(rvote) - 1;
    const uint32_t Lr = __shfl_sync(BK_FULL, L, row);
    const uint32_t before = __shfl_sync(BK_FULL, rincl - c, row);
    uint32_t k = rem - before;          // k-th set bit of Lr
    uint32_t bits = Lr;
    while (k--) bits &= bits - 1;
    const int col = __ffs(bits) - 1;
    const uint32_t id = uni((uint32_t)(o * 400 + row * 20 + col));    // same in all lanes -> uniform register
    o_out = (int)(id / 400u); row_out = (int)((id % 400u) / 20u); col_out = (int)(id % 20u);
    return id;
}

// Apply a placement for the side `player` -- engine/board.py:515-555 (cells, used, first-move flag,
// move_count, next player).  The frontier sets of the reference are not state here: A and C are
// recomputed from the cells each turn.
__device__ __forceinline__ void place_move(WarpState& s, uint32_t player, int o, int row, int col, int lane) {
    const uint32_t pat = c_orient_rowpat[o];                 // row k of the shape = bits 5k..5k+4
    const uint32_t k = (uint32_t)(lane - row);
    const uint32_t add = (k < 5u) ? (((pat >> (5u * k)) & 31u) << col) : 0u;
    const uint32_t pbit = 1u << (c_orient_piece[o] - 1);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (q == (int)player) { s.P[q] |= add; s.used[q] |= pbit; }
    }
    s.first_mask &= ~(1u << player);
    s.move_count += 1;
    s.to_move = (s.to_move + 1) & 3u;   // advances from current_player (engine/board.py:552,557-560)
}

// ---- scores ---------------------------------------------------------------------------------------
// Board.get_score (engine/board.py:562-577) and BlokusGame.get_score (engine/game.py:284-349).
__device__ __forceinline__ void scores_of(const WarpState& s, int lane, int32_t (&game)[4], int32_t (&board)[4]) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t row = s.P[q];
        uint32_t cells = __popc(row);
        uint32_t bonus = 0;
        if (lane == 0 || lane == 19) bonus += 5 * ((row & 1u) + ((row >> 19) & 1u));   // four board corners
        if (lane >= 8 && lane <= 11) bonus += 2 * __popc(row & 0xF00u);                 // rows/cols 8..11
        const uint32_t tot_cells = __reduce_add_sync(BK_FULL, cells);
        const uint32_t tot_bonus = __reduce_add_sync(BK_FULL, bonus);
        const uint32_t all = ((s.used[q] & BK_ALL_PIECES) == BK_ALL_PIECES) ? 15u : 0u;
        board[q] = (int32_t)(tot_cells + all);
        game[q] = (int32_t)(tot_cells + all + tot_bonus);
    }
}

// ---- Philox4x32-10 (same stream as oracle/blokus_oracle.c:orc_philox4x32_10) -------------------------
// Draw number t of (seed, g, stream): word t%4 of philox4x32_10(key = seed, ctr = (g_lo, g_hi, t/4, stream)).
struct PhiloxBlock { uint32_t w[4]; };
__device__ __forceinline__ PhiloxBlock philox_block(uint64_t seed, uint64_t g, uint32_t block, uint32_t stream) {
    uint32_t c0 = (uint32_t)g, c1 = (uint32_t)(g >> 32), c2 = block, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    PhiloxBlock b; b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
    return b;
}
__device__ __forceinline__ uint32_t philox_word(const PhiloxBlock& b, uint32_t t) {
    const uint32_t w = t & 3u;
    return w == 0 ? b.w[0] : w == 1 ? b.w[1] : w == 2 ? b.w[2] : b.w[3];
}
__device__ __forceinline__ uint32_t philox_draw(uint64_t seed, uint64_t g, uint32_t t, uint32_t stream) {
    uint32_t c0 = (uint32_t)g, c1 = (uint32_t)(g >> 32), c2 = t >> 2, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t w = t & 3u;
    return w == 0 ? c0 : w == 1 ? c1 : w == 2 ? c2 : c3;
}
