// snapshot.cuh -- win-probability snapshot feature rows on the device.
//
// The reference computes 34 features per (board, player) for its win-probability dataset and its learned MCTS leaf
// evaluator: analytics/winprob/features.py:23-58 (column order), :103-128 (board-level context), :131-193 (per
// player), with engine/advanced_metrics.py:8-15, :71-86, :89-105, :177-210 and engine/mobility_metrics.py:38-131.
// One call there costs 16 legal-move generations plus a BFS over the grid (~0.1 s in Python).  Here one warp owns one
// board: four piece-count generations (one per player, shared by the four feature rows), and the geometric features as
// row-sliced bit operations -- frontier sizes are popcounts, opponent adjacency is one neighbour mask, the dead-zone
// BFS is a flood fill over the empty cells that fills whole horizontal runs per step with a carry chain.
//
// The frontier sets of the reference are maintained incrementally and only by their owner's moves (engine/board.py:
// 315-367), so they keep cells that an opponent has occupied since: they are history, not a function of the cells.
// k_replay_snapshots therefore replays recorded games (move-id traces) with the same update rule and captures the rows
// where analytics/tournament/arena_runner.py:625-650,698 captures them; k_snapshot_features takes the sets as input (or
// rebuilds them from the cells, as a board restored from the wire has them).
//
// float64 arithmetic uses _rn intrinsics in the reference's operation order (no FMA contraction): rows are bit-identical.
#pragma once
#include "movegen_core.cuh"

#define BK_NFEAT 34
#define BK_SNAP_SMEM_PER_WARP (4 * 24 * 2 + 4 * BK_NFEAT * 8)     // piece counts + one staged feature block

// rows of a 4 x 7-word (400 bits per player) bitboard record; same layout as bk_state.bits
__device__ __forceinline__ void unpack_rows4(const uint64_t word, int lane, uint32_t (&rows)[4]) {
    const int bitpos = lane * 20;
    const int wi = bitpos >> 6, off = bitpos & 63;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint64_t lo = __shfl_sync(BK_FULL, word, (q * 7 + wi) & 31);
        uint64_t hi = __shfl_sync(BK_FULL, word, (q * 7 + wi + 1) & 31);
        uint64_t v = lo >> off;
        if (off) v |= hi << (64 - off);
        rows[q] = (lane < 20) ? ((uint32_t)v & BK_ROWMASK) : 0u;
    }
}
__device__ __forceinline__ void store_rows4(uint64_t* __restrict__ dst, int lane, const uint32_t (&rows)[4]) {
    uint64_t word = 0;
    const int q = lane / 7, w = lane - q * 7;
    const int r0 = (64 * w) / 20;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int rr = r0 + j;
        uint32_t v0 = __shfl_sync(BK_FULL, rows[0], rr & 31), v1 = __shfl_sync(BK_FULL, rows[1], rr & 31);
        uint32_t v2 = __shfl_sync(BK_FULL, rows[2], rr & 31), v3 = __shfl_sync(BK_FULL, rows[3], rr & 31);
        uint32_t v = (q == 0) ? v0 : (q == 1) ? v1 : (q == 2) ? v2 : v3;
        const int d = 20 * rr - 64 * w;
        if (rr < 20 && d < 64) word |= (d >= 0) ? ((uint64_t)v << d) : ((uint64_t)v >> (-d));
    }
    if (lane < 28) dst[lane] = word;
}

__device__ __forceinline__ uint32_t rows_orth(uint32_t x, int lane) {      // N/S/E/W neighbours of a row-sliced set
    uint32_t up = __shfl_up_sync(BK_FULL, x, 1), dn = __shfl_down_sync(BK_FULL, x, 1);
    if (lane == 0) up = 0;
    if (lane >= 19) dn = 0;
    return (((x << 1) | (x >> 1)) & BK_ROWMASK) | up | dn;
}
__device__ __forceinline__ uint32_t rows_diag(uint32_t x, int lane) {      // the four diagonal neighbours
    const uint32_t h = ((x << 1) | (x >> 1)) & BK_ROWMASK;
    uint32_t up = __shfl_up_sync(BK_FULL, h, 1), dn = __shfl_down_sync(BK_FULL, h, 1);
    if (lane == 0) up = 0;
    if (lane >= 19) dn = 0;
    return up | dn;
}

// Frontier sets of a board that carries no history: engine/board.py:261-313 (_compute_full_frontier), and the start
// corner while the first move is still to come (:389-405).
__device__ __forceinline__ void rebuild_frontiers(const WarpState& s, int lane, uint32_t (&F)[4]) {
    const uint32_t occ = s.P[0] | s.P[1] | s.P[2] | s.P[3];
    const uint32_t open = (lane < 20) ? (~occ & BK_ROWMASK) : 0u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        F[q] = open & rows_diag(s.P[q], lane) & ~rows_orth(s.P[q], lane);
        if ((s.first_mask >> q) & 1u) {
            const int crow = (q >= 2) ? 19 : 0;
            const uint32_t cbit = (q == 1 || q == 2) ? (1u << 19) : 1u;
            if (lane == crow) F[q] |= open & cbit;
        }
    }
}

// engine/board.py:315-367 after `player` placed the cells `add` (s already holds them): placed cells and their
// orthogonal neighbours leave the set; empty diagonal neighbours that touch no own cell by an edge join it.
__device__ __forceinline__ void update_frontier(const WarpState& s, uint32_t player, uint32_t add, int lane, uint32_t (&F)[4]) {
    const uint32_t occ = s.P[0] | s.P[1] | s.P[2] | s.P[3];
    const uint32_t own = sel4(s.P, player);
    const uint32_t gone = add | rows_orth(add, lane);
    const uint32_t join = rows_diag(add, lane) & ~occ & ~rows_orth(own, lane) & BK_ROWMASK;
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (q == (int)player) F[q] = (lane < 20) ? ((F[q] & ~gone) | join) : 0u;
}

// Empty cells 4-connected to a seed (advanced_metrics.py:177-210 runs a BFS per component).  Per step: every reached
// cell fills its whole horizontal run of empty cells (a carry chain towards the high bits, the same on the reversed
// row towards the low bits), then the rows above and below are entered.
__device__ __forceinline__ uint32_t flood_fill(uint32_t seeds, uint32_t open) {
    uint32_t reach = seeds & open;
    const uint32_t ropen = __brev(open);
    for (;;) {
        const uint32_t up = ((open + reach) ^ open) & open;
        const uint32_t rr = __brev(reach);
        const uint32_t dn = __brev(((ropen + rr) ^ ropen) & ropen);
        const uint32_t h = reach | up | dn;
        // lane 0 receives its own row from shfl_up and lanes >= 20 hold no open cells: no boundary fix-ups needed
        const uint32_t v = (__shfl_up_sync(BK_FULL, h, 1) | __shfl_down_sync(BK_FULL, h, 1)) & open;
        const uint32_t next = h | v;
        if (!__any_sync(BK_FULL, next != reach)) return next;
        reach = next;
    }
}

__device__ __forceinline__ int sel4i(const int (&v)[4], int i) {
    const int a = (i & 1) ? v[1] : v[0], b = (i & 1) ? v[3] : v[2];
    return (i & 2) ? b : a;
}

struct PieceTotalSink {            // legal placements per piece into shared memory
    static constexpr bool kRows = false;
    uint32_t acc = 0;
    uint16_t* out; int lane;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) { acc += __popc(L); }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {
        const uint32_t t = __reduce_add_sync(BK_FULL, acc);
        if (lane == 0) out[PID - 1] = (uint16_t)t;
        acc = 0;
    }
};

__device__ __noinline__ void piece_totals(const uint32_t A, const uint32_t C, const uint32_t used, uint16_t* out, int lane) {
    PieceTotalSink sink; sink.out = out; sink.lane = lane;
    bk_gen_all<PieceTotalSink, true>(A, C, used, sink);
}

// The 4 x 34 feature block of one board, left in shared memory.  pc = uint16[4][24] and feat = double[4][34] are this
// warp's shared scratch.
__device__ __forceinline__ void snapshot_features_block(const WarpState& s, const uint32_t (&F)[4], int turn_index, int max_turns,
                                                        int lane, uint16_t* pc, double* feat) {
    // ---- movegen-bound part: P_i for every player -------------------------------------------------------
    for (int k = lane; k < 4 * 24; k += 32) pc[k] = 0;
    __syncwarp();
#pragma unroll 1
    for (uint32_t p = 0; p < 4; p++) {
        uint32_t A, C; make_masks(s, p, lane, A, C);
        if (__any_sync(BK_FULL, C != 0)) piece_totals(A, C, sel4(s.used, p), pc + p * 24, lane);
    }
    __syncwarp();
    // ---- geometry, warp-wide ---------------------------------------------------------------------------------
    const uint32_t occ = s.P[0] | s.P[1] | s.P[2] | s.P[3];
    const uint32_t open = (lane < 20) ? (~occ & BK_ROWMASK) : 0u;
    const uint32_t reach = flood_fill(F[0] | F[1] | F[2] | F[3], open);
    const int dead = (int)__reduce_add_sync(BK_FULL, __popc(open & ~reach));
    const int occupied = (int)__reduce_add_sync(BK_FULL, __popc(occ));
    int fsize[4], cells[4], centre[4], adjacent[4];
    const int dr = (lane <= 9) ? 9 - lane : lane - 10;                       // advanced_metrics.py:78-80 in integers
#pragma unroll
    for (int q = 0; q < 4; q++) {
        fsize[q] = (int)__reduce_add_sync(BK_FULL, __popc(F[q]));
        cells[q] = (int)__reduce_add_sync(BK_FULL, __popc(s.P[q]));
        const uint32_t lo = s.P[q] & 0x3FFu, hi = s.P[q] >> 10;
        const int dc_lo = lo ? 9 - (31 - __clz(lo)) : 99, dc_hi = hi ? __ffs(hi) - 1 : 99;
        centre[q] = (int)__reduce_min_sync(BK_FULL, (uint32_t)min(9, dr + min(dc_lo, dc_hi)));
        adjacent[q] = (int)__reduce_add_sync(BK_FULL, __popc(s.P[q] & rows_orth(occ & ~s.P[q], lane)));
    }
    // ---- lane p assembles row p -----------------------------------------------------------------------------
    if (lane < 4) {
        const int p = lane;
        const uint32_t used = sel4(s.used, (uint32_t)p);
        int total = 0, used_count = 0, used_squares = 0, all_squares = 0;
        int by_size[6] = {0, 0, 0, 0, 0, 0};
        double on = 0.0, cw = 0.0, bucket[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int opp_moves = 0;
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            if (q == p) continue;
            for (int k = 0; k < BK_PIECES; k++) opp_moves += pc[q * 24 + k];
        }
#pragma unroll 1
        for (int pid = 1; pid <= BK_PIECES; pid++) {                          // mobility_metrics.py:85-104
            const int Oi = c_piece_first[pid + 1] - c_piece_first[pid];
            const int Si = c_orient_ncells[c_piece_first[pid]];
            all_squares += Si;
            if ((used >> (pid - 1)) & 1u) { used_count++; used_squares += Si; continue; }
            by_size[Si]++;
            const int Pi = pc[p * 24 + pid - 1];
            const double PNi = __ddiv_rn((double)Pi, (double)Oi);
            const double MWi = __dmul_rn(PNi, (double)Si);
            total += Pi;
            on = __dadd_rn(on, PNi);
            cw = __dadd_rn(cw, MWi);
            bucket[Si] = __dadd_rn(bucket[Si], MWi);
        }
        int opp_frontier = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) if (q != p) opp_frontier += fsize[q];
        const int my_f = sel4i(fsize, p);
        const double mt = (double)max(1, max_turns);
        double* f = feat + p * BK_NFEAT;
        f[0] = (double)my_f;
        f[1] = (double)(my_f - opp_frontier);
        f[2] = (double)sel4i(centre, p);
        f[3] = (double)sel4i(adjacent, p);
        f[4] = (double)dead;
        f[5] = __ddiv_rn((double)dead, 400.0);
        f[6] = (double)total;
        f[7] = on;
        f[8] = cw;
        f[9] = bucket[1]; f[10] = bucket[2]; f[11] = bucket[3]; f[12] = bucket[4]; f[13] = bucket[5];
        f[14] = (double)opp_moves;
        f[15] = __dadd_rn((double)my_f, cw);
        f[16] = (double)used_count;
        f[17] = (double)(BK_PIECES - used_count);
        f[18] = (double)(all_squares - used_squares);
        f[19] = (double)by_size[1]; f[20] = (double)by_size[2]; f[21] = (double)by_size[3];
        f[22] = (double)by_size[4]; f[23] = (double)by_size[5];
        f[24] = (double)(((used >> 16) & 1u) ^ 1u);
        f[25] = (double)(((used >> 18) & 1u) ^ 1u);
        f[26] = (double)(((used >> 19) & 1u) ^ 1u);
        f[27] = (double)s.move_count;
        f[28] = (double)turn_index;
        f[29] = __ddiv_rn((double)occupied, 400.0);
        f[30] = __ddiv_rn((double)used_count, 21.0);
        f[31] = __ddiv_rn((double)turn_index, mt);
        f[32] = __ddiv_rn((double)s.move_count, mt);
        f[33] = __ddiv_rn((double)sel4i(cells, p), 400.0);
    }
    __syncwarp();
}

// ... and written to `out` with coalesced stores.
__device__ __forceinline__ void snapshot_features(const WarpState& s, const uint32_t (&F)[4], int turn_index, int max_turns,
                                                  int lane, uint16_t* pc, double* feat, double* __restrict__ out) {
    snapshot_features_block(s, F, turn_index, max_turns, lane, pc, feat);
    for (int k = lane; k < 4 * BK_NFEAT; k += 32) out[k] = feat[k];
    __syncwarp();
}

// LearnedWinProbabilityEvaluator.predict_player_win_probability (mcts/learned_evaluator.py:133-143) for a
// "pairwise_logreg" model folded to 34 weights and a bias (model[0..33], model[34]): mean over the three opponents j of
// sigmoid(bias + sum_c w_c * (f_player[c] - f_j[c])), terms added in column order, opponents in player order.
__device__ __forceinline__ double win_probability(const double* feat, const double* __restrict__ model, uint32_t player, int lane) {
    double p = 0.0;
    if (lane < 3) {
        const uint32_t j = (uint32_t)lane + ((uint32_t)lane >= player ? 1u : 0u);          // the lane-th opponent
        double z = __ldg(model + BK_NFEAT);
        for (int c = 0; c < BK_NFEAT; c++)
            z = __dadd_rn(z, __dmul_rn(__ldg(model + c), __dsub_rn(feat[player * BK_NFEAT + c], feat[j * BK_NFEAT + c])));
        p = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-z)));
    }
    const double p0 = __shfl_sync(BK_FULL, p, 0), p1 = __shfl_sync(BK_FULL, p, 1), p2 = __shfl_sync(BK_FULL, p, 2);
    return __ddiv_rn(__dadd_rn(__dadd_rn(p0, p1), p2), 3.0);
}

// Feature rows of n boards.  frontiers: uint64[n][28] bitboards (or NULL: rebuilt from the cells); turn_index: int32[n]
// (or NULL: the board's move_count, as mcts/learned_evaluator.py:79-84 passes it).
__global__ void __launch_bounds__(256, 3)
k_snapshot_features(const bk_state* __restrict__ states, const uint64_t* __restrict__ frontiers,
                    const int32_t* __restrict__ turn_index, int n, int max_turns, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char snap_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double* feat = reinterpret_cast<double*>(snap_smem + wib * BK_SNAP_SMEM_PER_WARP);
    uint16_t* pc = reinterpret_cast<uint16_t*>(feat + 4 * BK_NFEAT);
    for (int i = warp; i < n; i += nwarps) {
        WarpState s; load_state(states + i, lane, s);
        uint32_t F[4];
        if (frontiers) unpack_rows4(lane < 28 ? __ldg(frontiers + (size_t)i * 28 + lane) : 0ull, lane, F);
        else rebuild_frontiers(s, lane, F);
        const int ti = turn_index ? turn_index[i] : (int)s.move_count;
        snapshot_features(s, F, ti, max_turns, lane, pc, feat, out + (size_t)i * 4 * BK_NFEAT);
    }
}

#define BK_MAX_CHECKPOINTS 32
struct bk_checkpoints { int32_t n; int32_t ply[BK_MAX_CHECKPOINTS]; };

// Replay recorded games and capture feature rows at the checkpoint plies: before the first turn (ply 0) and after every
// placement whose move_count is a checkpoint, with turn_index = the number of turns (passes included) played so far --
// the capture points of analytics/tournament/arena_runner.py:625-650,698.  A move that is not legal stops the replay of
// that game (its remaining checkpoints stay at turn_index -1).
__global__ void __launch_bounds__(256, 3)
k_replay_snapshots(const bk_state* __restrict__ start, const uint16_t* __restrict__ traces, int n, int n_turns,
                   const bk_checkpoints ck, int max_turns, double* __restrict__ features, int32_t* __restrict__ turn_index,
                   bk_state* __restrict__ snap_states, uint64_t* __restrict__ snap_frontiers) {
    extern __shared__ __align__(16) unsigned char snap_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double* feat = reinterpret_cast<double*>(snap_smem + wib * BK_SNAP_SMEM_PER_WARP);
    uint16_t* pc = reinterpret_cast<uint16_t*>(feat + 4 * BK_NFEAT);
    for (int g = warp; g < n; g += nwarps) {
        WarpState s;
        if (start) load_state(start + g, lane, s); else { init_empty(s); make_uniform(s); }
        uint32_t F[4]; rebuild_frontiers(s, lane, F);
        if (lane < ck.n) turn_index[(size_t)g * ck.n + lane] = -1;
        __syncwarp();
        const uint16_t* tr = traces + (size_t)g * n_turns;
        for (int t = -1; t < n_turns; t++) {
            if (t >= 0) {
                const uint32_t id = uni(tr[t]);
                if (id == BK_PASS) { s.to_move = (s.to_move + 1) & 3u; continue; }
                int o, row, col;
                const uint32_t p = s.to_move;
                if (!move_is_legal(s, p, id, lane, o, row, col)) break;
                const uint32_t before = sel4(s.P, p);
                place_move(s, p, o, row, col, lane);
                update_frontier(s, p, sel4(s.P, p) & ~before, lane, F);
            }
#pragma unroll 1
            for (int k = 0; k < ck.n; k++) {
                if (ck.ply[k] != (int)s.move_count) continue;
                const size_t slot = (size_t)g * ck.n + k;     // move_count only grows: each checkpoint matches once
                snapshot_features(s, F, t + 1, max_turns, lane, pc, feat, features + slot * 4 * BK_NFEAT);
                if (lane == 0) turn_index[slot] = t + 1;
                if (snap_states) store_state(snap_states + slot, lane, s);
                if (snap_frontiers) store_rows4(snap_frontiers + slot * 28, lane, F);
                __syncwarp();
            }
        }
    }
}

// The same for a "pairwise_gbt_phase" artifact (mcts/learned_evaluator.py:112-127) flattened by
// winprob.flatten_gbt_phase_models: the pair's phase (mean board occupancy < 0.33 / < 0.66 / else) picks the model, lane L
// walks trees L, L+32, ... (left when float32(x[feature]) <= threshold), the lane sums are combined with a butterfly of
// _rn additions (the oracle uses the same order), p = sigmoid(init + learning_rate * sum).
__device__ __forceinline__ double win_probability_gbt(const double* feat, const double* __restrict__ blob, uint32_t player, int lane) {
    double acc = 0.0;
#pragma unroll 1
    for (uint32_t n = 0; n < 3; n++) {
        const uint32_t j = n + (n >= player ? 1u : 0u);
        const double* fi = feat + player * BK_NFEAT;
        const double* fj = feat + j * BK_NFEAT;
        const double occ = __ddiv_rn(__dadd_rn(fi[29], fj[29]), 2.0);          // phase_board_occupancy of the pair
        const int phase = occ < 0.33 ? 0 : (occ < 0.66 ? 1 : 2);
        const double* m = blob + (int)__ldg(blob + phase);
        const int nt = (int)__ldg(m);
        const double init = __ldg(m + 1), lr = __ldg(m + 2);
        const double* roots = m + 4;
        const double* nodes = roots + nt;
        double partial = 0.0;
        for (int t = lane; t < nt; t += 32) {
            int k = (int)__ldg(roots + t);
            for (;;) {
                const double* rec = nodes + 5 * k;
                const int f = (int)__ldg(rec);
                if (f < 0) { partial = __dadd_rn(partial, __ldg(rec + 4)); break; }
                const double x = (double)(float)__dsub_rn(fi[f], fj[f]);
                k = (int)__ldg(rec + ((x <= __ldg(rec + 1)) ? 2 : 3));
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) partial = __dadd_rn(partial, __shfl_xor_sync(BK_FULL, partial, d));
        const double raw = __dadd_rn(init, __dmul_rn(lr, partial));
        const double p = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-raw)));
        acc = (n == 0) ? p : __dadd_rn(acc, p);
    }
    return __ddiv_rn(acc, 3.0);
}

// the evaluator's probability for either model form
__device__ __forceinline__ double win_probability_any(const double* feat, const double* __restrict__ model, int kind,
                                                      uint32_t player, int lane) {
    return kind == 1 ? win_probability_gbt(feat, model, player, lane) : win_probability(feat, model, player, lane);
}

// LearnedWinProbabilityEvaluator.potential (mcts/learned_evaluator.py:145-152): the clamped probability or its logit
__device__ __forceinline__ double potential_of(double p, int mode) {
    p = fmin(fmax(p, 1e-6), 1.0 - 1e-6);
    return mode == 1 ? log(__ddiv_rn(p, __dsub_rn(1.0, p))) : p;
}
