// heuristic.cuh -- HeuristicAgent as a device playout policy (agents/heuristic_agent.py:39-66, 222-244).
//
// select_action = softmax(T = 1) over the move scores, then one categorical draw.  The score of a move splits as
// score = s1 + s2 with s2 = 0.5*centre(anchor) and
//     s1 = size + 2*corners - 1.5*edge_score = SUM over the move's cells of (1 + 2*D(cell) - 1.5*f*B(cell)),
// D(cell) = diagonal neighbours of the cell inside the allowed mask A (corner_creation, :105-140), B(cell) = 1 within two
// cells of the border (:142-167), f = 1 early / 0.5 late (:169-176).  4*s1 is an integer in [-10, 188], so the softmax weight
// is e = exp(s1) * exp(s2) = g_e2[4*s1 + 10] * g_e1[anchor]: equal to the reference's exp(score - max)/sum up to the common
// factor and rounding; no max pass, no exp and no float64 score in the loop.  On the device, per turn:
//   1. the generator runs once; every lane (= board row) appends its NON-EMPTY rows of legal anchors to a private list in
//      shared memory (entry = orientation << 20 | row bits).  No zeroing, no vote, no pass over empty words.
//   2. one byte per cell v = 6 + 8*D - (6 early | 3 late)*B, so that 4*s1 + 10 = (sum of v over the cells) - 2*size + 10
//   3. the E entries, numbered lane by lane (row-major: row, orientation), are dealt out round-robin: lane j takes entries
//      j, j+32, ...  (rows near the player's corners hold most entries -- 30 of ~127 on average -- so a lane working through
//      its own list would leave the others idle).  The owner of entry g is found with a 5-step search over the lanes' list
//      bases (SHFL only).  Per move: five byte loads, two table loads, one product.
//   4. weights summed per lane in that order (columns ascending inside an entry), Hillis-Steele scan over lanes
//   5. target = total * r * 2^-32 with r the turn's Philox word; the lane whose inclusive sum first exceeds the target, then
//      the first move inside that lane's entries whose running sum exceeds it
// Every floating-point operation has a fixed order, restated in oracle/blokus_oracle.c (heuristic_pick), so the CPU oracle
// draws the same move.  The reference draws with numpy's RandomState.choice; as with the uniform policy the random stream is
// Philox (DESIGN.md "Random streams"), the distribution is the reference's.
// (Rounds 1-2 built an ascending move list first -- all 1 820 row words staged, prefix pass, two binary searches per move:
// 2 370 warp instructions per turn, 14 KB of shared memory per warp)
#pragma once
#include "movegen_core.cuh"

#define HEUR_PRIV_WORDS (BK_TBL_ORIENTS * 20)           /* entry i of lane r at word i*20 + r; a lane has <= 91 entries */
#define HEUR_E2_SIZE 200
/* per warp: private lists, allowed rows, cell bytes */
#define HEUR_SMEM_PER_WARP (HEUR_PRIV_WORDS * 4 + 32 * 4 + 400)
static_assert(HEUR_SMEM_PER_WARP % 16 == 0, "heuristic scratch keeps 16-byte alignment per warp");
// -DBK_HEUR_CHECK builds the policy with its own bounds checks (compute-sanitizer is not available on the GPU pool): every
// shared-memory and table index is tested and a violation traps the kernel (tools/heuristic_variants.py runs such a build).
#ifdef BK_HEUR_CHECK
#define HEUR_ASSERT(c) do { if (!(c)) __trap(); } while (0)
#else
#define HEUR_ASSERT(c) do { } while (0)
#endif

__device__ double g_e1[400];                 // exp(0.5 * (1 - dist(anchor, centre)/max_dist))   (heuristic_agent.py:185-198)
__device__ double g_e2[HEUR_E2_SIZE];        // exp((i - 10) / 4)
__device__ uint2 g_orient_off[BK_TBL_ORIENTS];   // x: cell offsets (row*20 + col) of cells 0..3, one byte each; y: cell 4 |
                                                 // (5 - size) << 8 | (10 - 2*size) << 16; short shapes repeat cell 0

__global__ void k_init_heuristic_tables() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 400) {
        const int ar = i / 20, ac = i - ar * 20;
        const double dr = __dsub_rn((double)ar, 9.5), dc = __dsub_rn((double)ac, 9.5);
        const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc)));
        const double maxd = __dsqrt_rn(__dmul_rn(2.0, __dmul_rn(9.5, 9.5)));
        g_e1[i] = exp(__dmul_rn(0.5, __dsub_rn(1.0, __ddiv_rn(dist, maxd))));
    }
    if (i < HEUR_E2_SIZE) g_e2[i] = exp(__dmul_rn((double)(i - 10), 0.25));
    if (i < BK_TBL_ORIENTS) {
        const uint32_t n = c_orient_ncells[i];
        uint32_t off[5];
        for (uint32_t k = 0; k < 5; k++) {
            const uint32_t rc = c_orient_cells[i * 5 + (k < n ? k : 0)];
            off[k] = (rc >> 4) * 20u + (rc & 15u);
        }
        g_orient_off[i] = make_uint2(off[0] | off[1] << 8 | off[2] << 16 | off[3] << 24, off[4] | (5u - n) << 8 | (10u - 2u * n) << 16);
    }
}

struct HeurScratch {
    uint32_t* priv;    // HEUR_PRIV_WORDS : lane-private lists of non-empty anchor rows
    uint32_t* arows;   // 32 : allowed rows A
    uint8_t* vmap;     // 400 : 6 + 8*D - k*B per cell
    __device__ __forceinline__ void bind(unsigned char* p) {
        priv = reinterpret_cast<uint32_t*>(p);
        arows = priv + HEUR_PRIV_WORDS;
        vmap = reinterpret_cast<uint8_t*>(arows + 32);
    }
};

// Takes the generator's rows as they come (kRows = false: from the mirrored generator bit j = column 19 - j).  The slot is
// kept as a 32-bit shared-memory address and the store is spelled out: as a C++ pointer the compiler carried a 64-bit
// generic copy along (IADD3 + IMAD.X per append next to the 32-bit one: 5 instructions instead of 3).
struct HeurPrivSink {
    static constexpr bool kRows = false;
    uint32_t a;        // shared-memory address of this lane's next free slot
#ifdef BK_HEUR_CHECK
    uint32_t a_end;
#endif
    template <int O, int SLOT, int GI> __device__ __forceinline__ void put(uint32_t L) {
        if (L) {
            HEUR_ASSERT(a < a_end && (L >> 20) == 0u);
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(L | ((uint32_t)O << 20)) : "memory");
            a += 80u;
        }
    }
    template <int G0, int NG, int PID> __device__ __forceinline__ void flush() {}
};

// softmax weight of the move of orientation table entry `tb` anchored at `cell`
__device__ __forceinline__ double heur_weight(const uint8_t* __restrict__ vmap, uint32_t cell, const uint2 tb) {
    const uint8_t* v = vmap + cell;
    HEUR_ASSERT(cell + (tb.x & 0xFFu) < 400u && cell + ((tb.x >> 8) & 0xFFu) < 400u && cell + ((tb.x >> 16) & 0xFFu) < 400u &&
                cell + (tb.x >> 24) < 400u && cell + (tb.y & 0xFFu) < 400u);
    const uint32_t v0 = v[tb.x & 0xFFu];
    const uint32_t sum = v0 + v[(tb.x >> 8) & 0xFFu] + v[(tb.x >> 16) & 0xFFu] + v[tb.x >> 24] + v[tb.y & 0xFFu];
    const uint32_t idx = sum + (tb.y >> 16) - ((tb.y >> 8) & 0xFFu) * v0;
    HEUR_ASSERT(idx < HEUR_E2_SIZE);
    return __dmul_rn(__ldg(g_e2 + idx), __ldg(g_e1 + cell));
}

// Returns 0 when `p` has no legal move; otherwise non-zero, and one move picked by the Philox word r comes back in
// id_out.  idx_out = its rank in the ascending id list, computed only when want_idx (the `choice` output of bk_playout).
// All lanes return the same values.
// `scratch_off` = byte offset of this warp's HEUR_SMEM_PER_WARP scratch inside the kernel's dynamic shared memory.  The
// pointers are rebuilt from the extern array HERE so that the compiler knows their address space: handed over as plain
// pointers they compiled to generic LD.E / ST.E with 64-bit address arithmetic, and those loads were the kernel's
// long-scoreboard stalls (profiles/r01r_all_kernels_ncu.md: 51 % issue-active, 1.8 long-scoreboard stalls per issue).
// Inlined: every kernel instantiation has exactly one call site, and out of line the callee spilled 88 bytes at the
// playout kernel's 80 registers and rebuilt the global-memory descriptors of its table loads (4 R2UR) for every move
// (6.6 -> 7.4 M playouts/s, 6.7 -> 7.5 M heuristic-rollout simulations/s when it was inlined).
__device__ __forceinline__ uint32_t heuristic_pick(const WarpState& s, uint32_t p, uint32_t r, uint32_t scratch_off, int lane,
                                                          bool want_idx, uint32_t& id_out, uint32_t& idx_out) {
    extern __shared__ __align__(16) unsigned char heur_dyn_smem[];
    HeurScratch sc; sc.bind(heur_dyn_smem + scratch_off);
    uint32_t A, C; make_masks(s, p, lane, A, C);
    __syncwarp();
    sc.arows[lane] = A;
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sc.priv + lane);
    HeurPrivSink sink; sink.a = a0;
#ifdef BK_HEUR_CHECK
    sink.a_end = a0 + 80u * BK_TBL_ORIENTS;
    HEUR_ASSERT(scratch_off % 16u == 0u);
#endif
    bk_gen_all<HeurPrivSink, true>(A, C, sel4(s.used, p), sink);
    const uint32_t cnt = (sink.a - a0) / 80u;                               // this lane's entries
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(BK_FULL, incl, d); if (lane >= d) incl += v; }
    const uint32_t E = __reduce_add_sync(BK_FULL, cnt);
    if (E == 0) return 0;
    const uint32_t base = incl - cnt;                                       // number of this lane's first entry
    __syncwarp();
    // ---- cell bytes v = 6 + 8*D - k*B of this lane's row; D[r][c] = A[r-1][c-1] + A[r-1][c+1] + A[r+1][c-1] + A[r+1][c+1]
    // as three bit planes (a 4-input adder), each plane's nibble spread to bytes with one multiply
    const uint32_t late = (__ddiv_rn((double)s.move_count, 100.0) < 0.3) ? 0u : 1u;      // heuristic_agent.py:169-172
    if (lane < 20) {
        const uint32_t up = lane > 0 ? sc.arows[lane - 1] : 0u, dn = lane < 19 ? sc.arows[lane + 1] : 0u;
        const uint32_t ul = up << 1, ur = up >> 1, dl = dn << 1, dr = dn >> 1;
        const uint32_t xa = ul ^ ur, ca = ul & ur, xb = dl ^ dr, cb = dl & dr, cc = xa & xb;
        const uint32_t p0 = xa ^ xb, p1 = ca ^ cb ^ cc, p2 = ca & cb;
        const uint32_t kk = late ? 3u : 6u;
        const bool edge_row = lane <= 2 || lane >= 17;
#pragma unroll
        for (int c4 = 0; c4 < 5; c4++) {
            const uint32_t bb = edge_row ? 0x01010101u : (c4 == 0 ? 0x00010101u : (c4 == 4 ? 0x01010100u : 0u));
            const uint32_t s0 = (((p0 >> (4 * c4)) & 15u) * 0x00204081u) & 0x01010101u;
            const uint32_t s1 = (((p1 >> (4 * c4)) & 15u) * 0x00204081u) & 0x01010101u;
            const uint32_t s2 = (((p2 >> (4 * c4)) & 15u) * 0x00204081u) & 0x01010101u;
            reinterpret_cast<uint32_t*>(sc.vmap)[lane * 5 + c4] = 0x06060606u - kk * bb + s0 * 8u + s1 * 16u + s2 * 32u;
        }
    }
    __syncwarp();
    // ---- per-lane sums of the weights: lane j works on entries j, j+32, ...
    const uint32_t trips = (E + 31u) >> 5;
    double sum = 0.0;
    for (uint32_t i = 0; i < trips; i++) {
        const uint32_t g = i * 32u + (uint32_t)lane;
        const uint32_t gg = g < E ? g : E - 1u;
        uint32_t src = 0;                                    // owner: the last lane whose base is <= gg
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) { const uint32_t b = __shfl_sync(BK_FULL, base, src + step); if (b <= gg) src += step; }
        HEUR_ASSERT(src < 20u && gg - __shfl_sync(BK_FULL, base, src) < __shfl_sync(BK_FULL, cnt, src));
        const uint32_t e = sc.priv[(gg - __shfl_sync(BK_FULL, base, src)) * 20u + src];
        HEUR_ASSERT((e >> 20) < BK_TBL_ORIENTS && (e & BK_ROWMASK) != 0u);
        uint32_t bits = g < E ? (e & BK_ROWMASK) : 0u;
        const uint2 tb = __ldg(g_orient_off + (e >> 20));
        const uint32_t cell0 = src * 20u + 19u;
        while (bits) {                                       // columns ascending = mirrored bits descending
            const uint32_t j = 31u - __clz(bits);
            bits ^= 1u << j;
            sum = __dadd_rn(sum, heur_weight(sc.vmap, cell0 - j, tb));
        }
    }
    double dincl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double v = __shfl_up_sync(BK_FULL, dincl, d); if (lane >= d) dincl = __dadd_rn(dincl, v); }
    const double total = __shfl_sync(BK_FULL, dincl, 31);
    const double target = __dmul_rn(total, __dmul_rn((double)r, 2.3283064365386963e-10));   // r * 2^-32
    const uint32_t vote = __ballot_sync(BK_FULL, dincl > target);
    const int Lp = vote ? (__ffs(vote) - 1) : 31;
    double acc = __shfl_up_sync(BK_FULL, dincl, 1);
    if (lane == 0) acc = 0.0;
    acc = __shfl_sync(BK_FULL, acc, Lp);
    // ---- walk lane Lp's entries (every lane does the same arithmetic, so the result is uniform without a broadcast)
    uint32_t pick_o = 0, pick_row = 0, pick_j = 0;
    bool found = false;
    for (uint32_t g = (uint32_t)Lp; g < E && !found; g += 32u) {
        uint32_t src = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) { const uint32_t b = __shfl_sync(BK_FULL, base, src + step); if (b <= g) src += step; }
        HEUR_ASSERT(src < 20u && g - __shfl_sync(BK_FULL, base, src) < __shfl_sync(BK_FULL, cnt, src));
        const uint32_t e = sc.priv[(g - __shfl_sync(BK_FULL, base, src)) * 20u + src];
        HEUR_ASSERT((e >> 20) < BK_TBL_ORIENTS && (e & BK_ROWMASK) != 0u);
        uint32_t bits = e & BK_ROWMASK;
        const uint2 tb = __ldg(g_orient_off + (e >> 20));
        pick_o = e >> 20; pick_row = src;
        while (bits) {
            const uint32_t j = 31u - __clz(bits);
            bits ^= 1u << j;
            pick_j = j;
            acc = __dadd_rn(acc, heur_weight(sc.vmap, src * 20u + 19u - j, tb));
            if (acc > target) { found = true; break; }
        }
    }
    // (if rounding kept the running sum at or below the target through the lane's last move, that last move is the pick;
    //  a lane whose inclusive sum exceeded the target owns at least one move)
    pick_o = uni(pick_o); pick_row = uni(pick_row); pick_j = uni(pick_j);
    id_out = pick_o * 400u + pick_row * 20u + (19u - pick_j);
    if (want_idx) {
        // rank in the ascending (orientation, row, column) list: every lane counts in its own list
        uint32_t below = 0;
        for (uint32_t i = 0; i < cnt; i++) {
            const uint32_t e = sc.priv[i * 20u + (uint32_t)lane];
            const uint32_t o = e >> 20, bits = e & BK_ROWMASK;
            if (o < pick_o || (o == pick_o && (uint32_t)lane < pick_row)) below += __popc(bits);
            else if (o == pick_o && (uint32_t)lane == pick_row) below += __popc(bits >> (pick_j + 1u));   // lower columns
        }
        idx_out = __reduce_add_sync(BK_FULL, below);
    }
    return E;
}
