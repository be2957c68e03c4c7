// heuristic.cuh -- HeuristicAgent as a device playout policy (agents/heuristic_agent.py:39-66, 222-244).
//
// select_action = softmax(T = 1) over the move scores, then one categorical draw.  On the device, per turn:
//   1. rows of all legal anchors -> shared memory (same generator, SmemRowSink)
//   2. ascending move list in shared memory (each lane owns 57 consecutive row words; one warp scan)
//   3. per move: corner / edge features (1 byte) and the softmax weight e_k = exp(s1_k) * exp(s2_k), where the score
//      is split as score = s1 + s2: s1 = size + 2*corners - 1.5*edge_score is a multiple of 0.25 with 1260 possible
//      values (table g_e2, indexed by the feature byte, the size and the game phase) and s2 = 0.5*centre(anchor)
//      depends on the anchor cell only (table g_e1).  exp(score - max)/sum of the reference equals e_k/sum e_k up to
//      rounding; no max pass, no exp and no float64 score in the loop (r01p profile: they were 26 % of the kernel)
//   4. e_k summed per lane (lane L owns moves k = L, L+32, ...), Hillis-Steele scan over lanes
//   5. target = total * r * 2^-32 with r the turn's Philox word; the lane whose inclusive sum first exceeds the
//      target, then the first move inside that lane whose running sum exceeds it
// Every floating-point operation has a fixed order, restated in oracle/blokus_oracle.c (heuristic_pick), so the
// CPU oracle draws the same move.  The reference draws with numpy's RandomState.choice; as with the uniform
// policy the random stream is Philox (DESIGN.md "Random streams"), the distribution is the reference's.
#pragma once
#include "movegen_core.cuh"

#define HEUR_WORDS_PER_LANE 57
#define HEUR_ROW_WORDS (32 * HEUR_WORDS_PER_LANE)       /* 1824 >= 91*20 */
#define HEUR_MAX_MOVES 2048
/* per warp: row words (bits | lane-local prefix << 20), ids, features, allowed rows, diagonal-count map, lane bases */
#define HEUR_SMEM_PER_WARP (HEUR_ROW_WORDS * 4 + HEUR_MAX_MOVES * 2 + HEUR_MAX_MOVES + 32 * 4 + 400 + 32 * 4)

__device__ double g_centre[400];      // 1 - dist(anchor, centre)/max_dist (heuristic_agent.py:185-198)
__device__ uint8_t g_edge[91 * 400];  // cells of move (o,row,col) within 2 cells of the border (heuristic_agent.py:153-167)
#define HEUR_E2_SIZE (2 * 6 * 21 * 5)
__device__ double g_e1[400];          // exp(0.5 * g_centre[anchor])
__device__ double g_e2[HEUR_E2_SIZE]; // exp(size + 2*corners - 1.5*edge_score), index ((late*6 + edge)*21 + corners)*5 + size-1

__global__ void k_init_heuristic_tables() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 400) {
        const int ar = i / 20, ac = i - ar * 20;
        const double dr = __dsub_rn((double)ar, 9.5), dc = __dsub_rn((double)ac, 9.5);
        const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc)));
        const double maxd = __dsqrt_rn(__dmul_rn(2.0, __dmul_rn(9.5, 9.5)));
        g_centre[i] = __dsub_rn(1.0, __ddiv_rn(dist, maxd));
        g_e1[i] = exp(__dmul_rn(0.5, g_centre[i]));
    }
    if (i < HEUR_E2_SIZE) {
        const int size = i % 5 + 1, corners = (i / 5) % 21, edge = (i / 105) % 6, late = i / 630;
        double s1 = __dadd_rn(0.0, __dmul_rn(1.0, (double)size));
        s1 = __dadd_rn(s1, __dmul_rn(2.0, (double)corners));
        const double edge_score = late ? __dmul_rn((double)edge, 0.5) : (double)edge;
        g_e2[i] = exp(__dadd_rn(s1, __dmul_rn(-1.5, edge_score)));
    }
    if (i < 91 * 400) {
        const int o = i / 400, cell = i - o * 400, ar = cell / 20, ac = cell - ar * 20;
        int edge = 0;
        for (int k = 0; k < c_orient_ncells[o]; k++) {
            const uint32_t rc = c_orient_cells[o * 5 + k];
            const int r = ar + (int)(rc >> 4), c = ac + (int)(rc & 15);
            edge += (min(min(r, c), min(19 - r, 19 - c)) <= 2);
        }
        g_edge[i] = (uint8_t)edge;
    }
}

struct HeurScratch {
    uint32_t* rows;    // HEUR_ROW_WORDS : bits 0..19 legal anchors of (orientation,row) word; bits 20..30 lane-local prefix
    uint16_t* ids;     // HEUR_MAX_MOVES : ascending move ids
    uint8_t* feat;     // HEUR_MAX_MOVES : corners | edge << 5
    uint32_t* arows;   // 32 : allowed rows A
    uint8_t* dmap;     // 400 : number of diagonal neighbours of each cell that lie in A (0..4)
    uint32_t* base;    // 32 : first list position owned by each lane's 57 row words
    __device__ __forceinline__ void bind(unsigned char* p) {
        rows = reinterpret_cast<uint32_t*>(p);
        ids = reinterpret_cast<uint16_t*>(rows + HEUR_ROW_WORDS);
        feat = reinterpret_cast<uint8_t*>(ids + HEUR_MAX_MOVES);
        arows = reinterpret_cast<uint32_t*>(feat + HEUR_MAX_MOVES);
        dmap = reinterpret_cast<uint8_t*>(arows + 32);
        base = reinterpret_cast<uint32_t*>(dmap + 400);
    }
};

struct HeurRowSink {
    static constexpr bool kRows = true;
    uint32_t* rows; int lane;
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) { if (lane < 20) rows[O * 20 + lane] = L; }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
};

// corner_creation of a move = sum over its cells of the diagonal-neighbour count map; edge from the static table
__device__ __forceinline__ uint32_t heur_features(const uint8_t* __restrict__ dmap, uint32_t id) {
    const uint32_t o = id / 400u, cell = id - o * 400u;
    const uint32_t cells = c_orient_cells32[o];          // 5 x (row<<3 | col), padded with repeats of cell 0
    const uint32_t n = c_orient_ncells[o];
    uint32_t corners = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint32_t rc = (cells >> (6 * i)) & 63u;
        const uint32_t d = dmap[cell + (rc >> 3) * 20u + (rc & 7u)];
        corners += (i < (int)n) ? d : 0u;
    }
    return corners | ((uint32_t)g_edge[id] << 5);
}

// softmax weight of a move from its stored feature byte (corners | edge << 5): two table loads and one product
__device__ __forceinline__ double heur_weight(uint32_t feat, uint32_t id, uint32_t late) {
    const uint32_t o = id / 400u, cell = id - o * 400u;
    const uint32_t idx = ((late * 6u + (feat >> 5)) * 21u + (feat & 31u)) * 5u + (c_orient_ncells[o] - 1u);
    return __dmul_rn(__ldg(g_e2 + idx), __ldg(g_e1 + cell));
}

// Returns the number of legal moves of `p`; when > 0 picks one by the Philox word r and returns it through
// (id, idx = rank in the ascending list).  All lanes return the same values.
// `scratch_off` = byte offset of this warp's HEUR_SMEM_PER_WARP scratch inside the kernel's dynamic shared memory.  The
// pointers are rebuilt from the extern array HERE so that the compiler knows their address space: handed over as plain
// pointers they compiled to generic LD.E / ST.E with 64-bit address arithmetic, and those loads were the kernel's
// long-scoreboard stalls (profiles/r01r_all_kernels_ncu.md: 51 % issue-active, 1.8 long-scoreboard stalls per issue).
__device__ __noinline__ uint32_t heuristic_pick(const WarpState s, uint32_t p, uint32_t r, uint32_t scratch_off, int lane,
                                                uint32_t& id_out, uint32_t& idx_out) {
    extern __shared__ __align__(16) unsigned char heur_dyn_smem[];
    HeurScratch sc; sc.bind(heur_dyn_smem + scratch_off);
    uint32_t A, C; make_masks(s, p, lane, A, C);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < HEUR_WORDS_PER_LANE; k++) sc.rows[k * 32 + lane] = 0;
    sc.arows[lane] = A;
    __syncwarp();
    HeurRowSink sink; sink.rows = sc.rows; sink.lane = lane;
    bk_gen_all(A, C, sel4(s.used, p), sink);
    __syncwarp();
    // ---- diagonal-count map of A (this lane's row): D[r][c] = A[r-1][c-1] + A[r-1][c+1] + A[r+1][c-1] + A[r+1][c+1]
    if (lane < 20) {
        const uint32_t up = lane > 0 ? sc.arows[lane - 1] : 0u, dn = lane < 19 ? sc.arows[lane + 1] : 0u;
        const uint32_t ul = up << 1, ur = up >> 1, dl = dn << 1, dr = dn >> 1;
#pragma unroll
        for (int c4 = 0; c4 < 5; c4++) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 4 * c4 + j;
                w |= (((ul >> c) & 1u) + ((ur >> c) & 1u) + ((dl >> c) & 1u) + ((dr >> c) & 1u)) << (8 * j);
            }
            reinterpret_cast<uint32_t*>(sc.dmap)[lane * 5 + c4] = w;
        }
    }
    // ---- ascending move list: lane L owns row words [57L, 57L+57); lane-local prefixes go into bits 20..30 of
    // the words, lane bases into sc.base, and every list position k is then located by two binary searches, so
    // all lanes extract the same number of moves (no divergent per-word bit loops)
    uint32_t* mine = sc.rows + lane * HEUR_WORDS_PER_LANE;
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < HEUR_WORDS_PER_LANE; k++) {
        const uint32_t bits = mine[k];
        mine[k] = bits | (cnt << 20);
        cnt += __popc(bits);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(BK_FULL, incl, d); if (lane >= d) incl += v; }
    uint32_t n = __reduce_add_sync(BK_FULL, cnt);
    if (n == 0) return 0;
    if (n > HEUR_MAX_MOVES) n = HEUR_MAX_MOVES;            // never reached in play (largest list seen: 1003)
    sc.base[lane] = incl - cnt;
    __syncwarp();
    for (uint32_t k = lane; k < n; k += 32) {
        uint32_t L = 0;                                     // owner lane: the last one whose base is <= k
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) if (sc.base[L + step] <= k) L += step;
        const uint32_t rem = k - sc.base[L];
        const uint32_t* words = sc.rows + L * HEUR_WORDS_PER_LANE;
        uint32_t w = 0;                                     // last word whose prefix is <= rem ...
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1) if (w + step < HEUR_WORDS_PER_LANE && (words[w + step] >> 20) <= rem) w += step;
        uint32_t word = words[w];
        // ... and which actually holds the move: words with equal prefix before it are empty, so the LAST such word
        // is the first non-empty one at that prefix only if no later word shares it -- the search lands on the last
        // word with prefix <= rem, whose bit count covers rem by construction
        uint32_t bits = word & BK_ROWMASK;
        for (uint32_t skip = rem - (word >> 20); skip; skip--) bits &= bits - 1;
        sc.ids[k] = (uint16_t)((L * HEUR_WORDS_PER_LANE + w) * 20u + (__ffs(bits) - 1));
    }
    __syncwarp();
    // ---- features and per-lane sums of the weights
    const uint32_t late = (__ddiv_rn((double)s.move_count, 100.0) < 0.3) ? 0u : 1u;      // heuristic_agent.py:169-172
    double sum = 0.0;
    for (uint32_t k = lane; k < n; k += 32) {
        const uint32_t id = sc.ids[k];
        const uint32_t f = heur_features(sc.dmap, id);
        sc.feat[k] = (uint8_t)f;
        sum = __dadd_rn(sum, heur_weight(f, id, late));
    }
    __syncwarp();
    double dincl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double v = __shfl_up_sync(BK_FULL, dincl, d); if (lane >= d) dincl = __dadd_rn(dincl, v); }
    const double total = __shfl_sync(BK_FULL, dincl, 31);
    const double target = __dmul_rn(total, __dmul_rn((double)r, 2.3283064365386963e-10));   // r * 2^-32
    const uint32_t vote = __ballot_sync(BK_FULL, dincl > target);
    const int L = vote ? (__ffs(vote) - 1) : 31;
    double acc = __shfl_up_sync(BK_FULL, dincl, 1);
    if (lane == 0) acc = 0.0;
    acc = __shfl_sync(BK_FULL, acc, L);
    // walk lane L's moves (every lane does the same arithmetic, so the result is uniform without a broadcast)
    uint32_t pick = (uint32_t)L;
    for (uint32_t k = (uint32_t)L; k < n; k += 32) {
        pick = k;
        acc = __dadd_rn(acc, heur_weight(sc.feat[k], sc.ids[k], late));
        if (acc > target) break;
    }
    if (pick >= n) pick = n - 1;     // only if lane L owns no move at all (cannot happen: its sum exceeded target)
    idx_out = uni(pick);
    id_out = uni((uint32_t)sc.ids[pick]);
    return n;
}
