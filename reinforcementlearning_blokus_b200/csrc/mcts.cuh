// mcts.cuh -- K5: UCT search with device rollouts: one warp per tree (root-parallel across warps/GPUs), or a CTA of up
// to 16 warps per tree that batches the rollouts of every new leaf (leaf_batch, see k_mcts).
//
// Mirrors MCTSAgent of the reference (mcts/mcts_agent.py):
//   iteration (:360-382)  select -> expand one child -> simulate -> back-propagate the same reward to the root
//   selection (:384-406)  descend while the node has no untried move and has children; child = first argmax of
//                         Q/n + c*sqrt(ln N_parent / n)   (:68-89; FastMCTS variant c*sqrt(2 ln N / n), fast_mcts_agent.py:45-53)
//   expansion (:113-145)  pops the LAST untried move; untried moves are the node's legal moves in ascending id
//                         order here, so the k-th expansion of a node takes legal move number n_legal-1-k
//   simulation (:408-437) transposition-table probe on the full-board Zobrist key, else rollout and store
//   rollout (:470-554)    at most max_rollout_moves plies; stops at the first player without a move (no passes);
//                         reward = Board.get_score(node player) after - before
//   result (:180-191)     visits / reward sums of the root's children
// A tree never stores move lists: a node keeps n_legal and n_expanded, and the wanted move is re-derived with
// the row-sliced generator (count per orientation -> orientation -> row -> column).
#pragma once
#include "movegen_core.cuh"
#include "heuristic.cuh"
#include "snapshot.cuh"
// (included by blokus_b200.cu after zobrist_of / tt_lookup / tt_insert / move_is_legal are defined)

struct MctsLayout {            // byte offsets inside one tree's workspace slice
    uint32_t max_nodes, pool_slots;
    size_t off_states, off_parent, off_visits, off_reward, off_child_base, off_n_legal, off_n_expanded,
           off_move, off_pool, off_header, off_frontier, off_value, off_prior, total;
};

__host__ __device__ inline size_t bk_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline MctsLayout mcts_layout(int iterations_total, bool with_frontier = false, int max_depth = 0) {
    MctsLayout L;
    size_t T = (size_t)(iterations_total > 0 ? iterations_total : 1);
    // a flat bandit (max_depth 1) never holds more than the root and one node per root move, however long it runs
    if (max_depth == 1 && T > 2048) T = 2048;
    L.max_nodes = (uint32_t)(T + 1);
    // a node first expanded at iteration t can still gain at most T - t children, and never more than its number
    // of legal moves (largest list seen in 10^6 states: 1003); a tree that outgrows the pool reports status 1
    size_t pool = T * (T + 1) / 2 + 1;
    if (pool > T * 2048) pool = T * 2048;
    if (max_depth == 1 && pool > 2049) pool = 2049;          // only the root has children
    L.pool_slots = (uint32_t)(pool > 0xFFFFFF00ull ? 0xFFFFFF00ull : pool);
    size_t o = 0;
    L.off_states = o;      o += (size_t)L.max_nodes * sizeof(bk_state);
    L.off_reward = o;      o += (size_t)L.max_nodes * 8;
    L.off_parent = o;      o += (size_t)L.max_nodes * 4;
    L.off_visits = o;      o += (size_t)L.max_nodes * 4;
    L.off_child_base = o;  o += (size_t)L.max_nodes * 4;
    L.off_pool = o;        o += (size_t)L.pool_slots * 4;
    L.off_n_legal = o;     o += (size_t)L.max_nodes * 2;
    L.off_n_expanded = o;  o += (size_t)L.max_nodes * 2;
    L.off_move = o;        o += (size_t)L.max_nodes * 2;
    o = bk_align_up(o, 8);
    L.off_header = o;      o += 128;
    // learned leaf evaluation: every node also keeps the four frontier sets of its board (28 words, engine/board.py:315-367)
    L.off_frontier = o;    if (with_frontier) o += (size_t)L.max_nodes * 28 * 8;
    // ... its own learned value V(board, player to move) and its progressive-bias prior (mcts_agent.py:453-468)
    L.off_value = o;       if (with_frontier) o += (size_t)L.max_nodes * 8;
    L.off_prior = o;       if (with_frontier) o += (size_t)L.max_nodes * 8;
    L.total = bk_align_up(o, 256);
    return L;
}

#define BK_MCTS_MAX_WARPS 16     /* leaf-batched rollouts: warps per tree */
struct MctsHeader {            // persistent between launches of the same search (resume); 92 of the slot's 128 bytes
    uint32_t n_nodes, pool_used, iterations_done, draws, overflow, root_children, root_player;
    int32_t root_score;
    uint32_t draws_x[BK_MCTS_MAX_WARPS - 1];   // draw counters of warps 1..15 (leaf-batched rollouts)
};

static_assert(sizeof(MctsHeader) <= 128, "MctsHeader must fit its workspace slot");

struct MctsTT {                // optional transposition table (mcts/zobrist.py:155-220); keys == nullptr: off
    unsigned long long* keys; double* rewards; unsigned long long* occupied; uint32_t* stamps; uint32_t log2_slots;
    const uint64_t* zobrist;   // 2088-entry table, reference order
};

// The group slots (one row per warp of the CTA) are this function's own static shared array (a pointer parameter would hide the
// address space and turn the sink's STS into generic stores).
template <int NW>
__device__ __noinline__ Totals gen_totals(uint32_t A, uint32_t C, uint32_t used, int lane) {
    __shared__ __align__(16) uint32_t slots[NW][BK_GROUP_SLOTS];
    GroupSink sink; sink.begin(slots[threadIdx.x >> 5], lane);
    bk_gen_all<GroupSink, true>(A, C, used, sink);
    return reduce_totals(sink, lane);
}

__device__ __forceinline__ int32_t board_score_of(const WarpState& s, uint32_t q) {
    const uint32_t cells = __reduce_add_sync(BK_FULL, (uint32_t)__popc(sel4(s.P, q)));
    return (int32_t)(cells + (((sel4(s.used, q) & BK_ALL_PIECES) == BK_ALL_PIECES) ? 15u : 0u));
}

#define BK_MCTS_THREADS 32   /* one warp = one tree per CTA, so a handful of trees still spreads over the SMs */
#ifndef BK_MCTS_CTAS_PER_SM
#define BK_MCTS_CTAS_PER_SM 24
#endif
// heuristic rollouts: 7.8 KB of list scratch per warp (round 1: 14 KB, which capped this at 14); measured on B200 with the
// policy inlined: 6.9 M simulations/s at 16 one-warp CTAs per SM, 7.5 M at 20 (96 registers), 7.5 M at 24 (80 registers, spills)
#ifndef BK_MCTS_HEUR_CTAS_PER_SM
#define BK_MCTS_HEUR_CTAS_PER_SM 20
#endif

// BATCH = false: one warp per tree (CTA of 32 threads), the reference's one-rollout-per-iteration search.
// BATCH = true : "leaf-batched rollouts" -- a CTA of W <= 16 warps owns the tree; warp 0 selects and expands exactly as
// before, then ALL W warps roll out from the new leaf with their own Philox streams (stream 1 + w of the tree) and the
// W rewards are backed up together (visits += W).  One tree then advances ~W simulations per iteration latency, which
// is what a single agent move needs (BASELINE config 4); W = 1 is bit-identical to BATCH = false.
// LEARNED = true: learned leaf evaluation in place of the rollout (mcts_agent.py:427-451) with a "pairwise_logreg"
// win-probability model folded to 34 weights + bias (prm.leaf_model): the leaf's 4 x 34 snapshot features are computed in
// shared memory (snapshot.cuh: four piece-count generations, flood fill) and the reward is the mean pairwise
// probability of the player to move at the leaf.  Rewards are doubles then (same 8-byte slots), and every node carries
// the frontier sets of its board, updated on expansion with the reference's rule.
template <int POLICY, bool BATCH = false, bool LEARNED = false>
__global__ void __launch_bounds__(BATCH ? 32 * BK_MCTS_MAX_WARPS : BK_MCTS_THREADS,
                                  BATCH ? 1 : (LEARNED ? 16 : (POLICY == 0 ? BK_MCTS_CTAS_PER_SM : BK_MCTS_HEUR_CTAS_PER_SM)))
k_mcts(unsigned int* __restrict__ ticket, const bk_state* __restrict__ roots, int n_roots, int n_trees, bk_mcts_params prm, MctsTT tt,
       unsigned char* __restrict__ workspace, int resume, int max_root_moves,
       int32_t* __restrict__ out_visits, long long* __restrict__ out_reward, uint32_t* __restrict__ out_counts,
       uint16_t* __restrict__ out_moves, uint32_t* __restrict__ out_status) {
    extern __shared__ __align__(16) unsigned char heur_smem[];    // HEUR_SMEM_PER_WARP bytes per warp when rollout_policy == 1
    __shared__ int sh_tree;
    __shared__ uint32_t sh_node, sh_stop, sh_cached, sh_root_player;
    __shared__ int32_t sh_root_score;
    __shared__ long long sh_r[BK_MCTS_MAX_WARPS];
    __shared__ __align__(16) double s_feat[LEARNED ? 4 * BK_NFEAT : 1];      // leaf feature block (learned evaluation)
    __shared__ uint16_t s_pc[LEARNED ? 4 * 24 : 1];
    const int lane = threadIdx.x & 31;
    const int wib = BATCH ? (int)(threadIdx.x >> 5) : 0;          // warp in block; warp 0 leads the tree
    const int W = BATCH ? (int)(blockDim.x >> 5) : 1;
    const bool lead = wib == 0;
    const uint32_t hs = (uint32_t)wib * HEUR_SMEM_PER_WARP;       // this warp's heuristic scratch
    const MctsLayout L = mcts_layout(prm.iterations_total, LEARNED, prm.max_depth);
    // trees differ in cost (rollout lengths, tree depth): the next tree comes from a per-launch ticket counter so that
    // a batch larger than the resident CTAs has no idle tail; a tree's result depends on (seed, tree id) only
    for (;;) {
        int tree;
        {
            uint32_t t0 = 0;
            if (lead && lane == 0) t0 = atomicAdd(ticket, 1u);
            tree = (int)uni(t0);
            if (BATCH) {
                if (lead && lane == 0) sh_tree = tree;
                __syncthreads();
                tree = sh_tree;
                __syncthreads();
            }
        }
        if (tree >= n_trees) break;
        unsigned char* base = workspace + (size_t)tree * L.total;
        bk_state* states = reinterpret_cast<bk_state*>(base + L.off_states);
        long long* reward = reinterpret_cast<long long*>(base + L.off_reward);
        int32_t* parent = reinterpret_cast<int32_t*>(base + L.off_parent);
        int32_t* visits = reinterpret_cast<int32_t*>(base + L.off_visits);
        uint32_t* child_base = reinterpret_cast<uint32_t*>(base + L.off_child_base);
        uint32_t* pool = reinterpret_cast<uint32_t*>(base + L.off_pool);
        uint16_t* n_legal = reinterpret_cast<uint16_t*>(base + L.off_n_legal);
        uint16_t* n_expanded = reinterpret_cast<uint16_t*>(base + L.off_n_expanded);
        uint16_t* move_of = reinterpret_cast<uint16_t*>(base + L.off_move);
        MctsHeader* hdr = reinterpret_cast<MctsHeader*>(base + L.off_header);
        uint64_t* frontier = reinterpret_cast<uint64_t*>(base + L.off_frontier);       // [node][28], LEARNED only
        double* reward_d = reinterpret_cast<double*>(reward);                          // the same slots as doubles
        double* node_value = reinterpret_cast<double*>(base + L.off_value);            // LEARNED only
        double* node_prior = reinterpret_cast<double*>(base + L.off_prior);
        const bool bias_on = LEARNED && prm.progressive_bias_weight != 0.0;
        uint32_t F[4] = {0u, 0u, 0u, 0u};                                               // frontier rows of `s`
        const uint64_t tree_id = prm.tree_offset + (uint64_t)tree;
        // Transposition-table namespace: tree k of a launch reads and writes keys salted with tt_tree_base + k, so trees
        // searched concurrently never see each other's entries (a tree's result depends on (seed, tree id, table
        // contents of its namespace) only).  The namespace does NOT follow tree_offset: an agent keeps it constant over
        // its searches, and search k+1 finds what search k cached (mcts_agent.py:337-339, 418-434).  Namespace 0 uses
        // the plain Zobrist key -- the reference's one table.
        const uint64_t salt = (uint64_t)(uint32_t)(prm.tt_tree_base + tree) * 0xD6E8FEB86659FD93ull;
        uint32_t tt_hits = 0, tt_probes = 0;

        uint32_t n_nodes = 0, pool_used = 0, iter_done = 0, draws = 0, overflow = 0, root_player = 0;
        int32_t root_score = 0;
        WarpState s;
        if (lead) {
            if (!resume) {
                load_state(roots + (tree % n_roots), lane, s);
                uint32_t A, C; make_masks(s, s.to_move, lane, A, C);
                const Totals t0 = gen_totals<BATCH ? BK_MCTS_MAX_WARPS : 1>(A, C, sel4(s.used, s.to_move), lane);
                store_state(states + 0, lane, s);
                if (LEARNED) {
                    if (prm.root_frontiers)
                        unpack_rows4(lane < 28 ? __ldg(prm.root_frontiers + (size_t)(tree % n_roots) * 28 + lane) : 0ull, lane, F);
                    else
                        rebuild_frontiers(s, lane, F);
                    store_rows4(frontier, lane, F);
                    if (bias_on) {                 // V(root board, root player): the parent value of the root's children
                        snapshot_features_block(s, F, (int)s.move_count, prm.leaf_max_turns, lane, s_pc, s_feat);
                        const double v0 = win_probability_any(s_feat, prm.leaf_model, prm.leaf_model_kind, s.to_move, lane);
                        if (lane == 0) { node_value[0] = v0; node_prior[0] = 0.0; }
                        __syncwarp();
                    }
                }
                if (lane == 0) {
                    parent[0] = -1; visits[0] = 0; reward[0] = 0; child_base[0] = 0xFFFFFFFFu;
                    n_legal[0] = (uint16_t)t0.n; n_expanded[0] = 0; move_of[0] = 0xFFFF;
                }
                n_nodes = 1; pool_used = 0; iter_done = 0; draws = 0;
                root_player = s.to_move; root_score = board_score_of(s, s.to_move);
                __syncwarp();
            } else {
                n_nodes = hdr->n_nodes; pool_used = hdr->pool_used; iter_done = hdr->iterations_done;
                draws = hdr->draws; overflow = hdr->overflow; root_player = hdr->root_player; root_score = hdr->root_score;
            }
            if (BATCH && lane == 0) { sh_root_player = root_player; sh_root_score = root_score; }
        } else if (resume) {
            draws = hdr->draws_x[wib - 1];
        }
        if (BATCH) {
            __syncthreads();
            root_player = sh_root_player; root_score = sh_root_score;
        }

        // rollout draws: 4 per Philox block (draw t = word t%4 of block t/4 of this warp's stream 1 + w); the block is
        // kept in registers and recomputed when the draw counter enters the next one
        const uint32_t stream = 1u + (uint32_t)wib;
        PhiloxBlock blk = philox_block(prm.seed, tree_id, draws >> 2, stream);
        uint32_t blk_of = draws >> 2;
        auto next_draw = [&]() -> uint32_t {
            if ((draws >> 2) != blk_of) { blk_of = draws >> 2; blk = philox_block(prm.seed, tree_id, blk_of, stream); }
            return philox_word(blk, draws++);
        };
        for (int it = 0; it < prm.iterations; it++) {
            uint32_t node = 0, depth = 0;
            bool have_value = false;               // learned value of the new leaf already computed during expansion
            double leaf_value = 0.0;
            if (lead && !overflow) {
            // ---- selection (mcts_agent.py:384-406)
            while (true) {
                const uint32_t nl = n_legal[node], ne = n_expanded[node];
                const bool may_expand = ne < nl && (prm.max_depth <= 0 || (int)depth < prm.max_depth);
                if (may_expand || ne == 0) break;
                const double logN = log((double)visits[node]) * (prm.ucb_variant ? 2.0 : 1.0);
                // reward_mode 2 counts rewards in halves (the constant 0.5 of the reference's dummy evaluator); x1.0 is exact
                const double unit = prm.reward_mode == 2 ? 0.5 : 1.0;
                const uint32_t cb = child_base[node];
                double best = -1.0; uint32_t best_k = 0xFFFFFFFFu; bool have = false;
                for (uint32_t k = lane; k < ne; k += 32) {
                    const uint32_t ch = pool[cb + k];
                    const double v = (double)visits[ch];
                    double u;
                    if (visits[ch] == 0) u = __longlong_as_double(0x7FF0000000000000ll);
                    else {
                        u = __dadd_rn(__ddiv_rn(LEARNED ? reward_d[ch] : __dmul_rn((double)reward[ch], unit), v),
                                      __dmul_rn(prm.exploration, sqrt(__ddiv_rn(logN, v))));
                        if (bias_on)               // + w * (prior / (1 + n)), mcts_agent.py:88-89
                            u = __dadd_rn(u, __dmul_rn(prm.progressive_bias_weight, __ddiv_rn(node_prior[ch], __dadd_rn(1.0, v))));
                    }
                    if (!have || u > best) { best = u; best_k = k; have = true; }
                }
                // warp argmax, ties -> smallest k (Python's max keeps the first maximal child)
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const double ob = __shfl_xor_sync(BK_FULL, best, d);
                    const uint32_t ok = __shfl_xor_sync(BK_FULL, best_k, d);
                    const bool oh = ok != 0xFFFFFFFFu;
                    if (oh && (best_k == 0xFFFFFFFFu || ob > best || (ob == best && ok < best_k))) { best = ob; best_k = ok; }
                }
                node = pool[cb + best_k];
                depth++;
            }
            // ---- expansion (mcts_agent.py:113-145): pop the last untried move of `node`
            load_state(states + node, lane, s);
            if (LEARNED) unpack_rows4(lane < 28 ? frontier[(size_t)node * 28 + lane] : 0ull, lane, F);
            {
                const uint32_t nl = n_legal[node], ne = n_expanded[node];
                if (ne < nl && (prm.max_depth <= 0 || (int)depth < prm.max_depth)) {
                    const uint32_t p = s.to_move;
                    uint32_t A, C; make_masks(s, p, lane, A, C);
                    const Totals tot = gen_totals<BATCH ? BK_MCTS_MAX_WARPS : 1>(A, C, sel4(s.used, p), lane);
                    int o, row, col;
                    const uint32_t id = select_move(tot, nl - 1 - ne, A, C, lane, o, row, col);
                    const uint32_t own_before = sel4(s.P, p);
                    place_move(s, p, o, row, col, lane);
                    if (LEARNED) update_frontier(s, p, sel4(s.P, p) & ~own_before, lane, F);
                    uint32_t A2, C2; make_masks(s, s.to_move, lane, A2, C2);
                    const Totals tc = gen_totals<BATCH ? BK_MCTS_MAX_WARPS : 1>(A2, C2, sel4(s.used, s.to_move), lane);
                    uint32_t cb = child_base[node];
                    if (cb == 0xFFFFFFFFu) {
                        const uint32_t remaining = (uint32_t)prm.iterations_total - iter_done;
                        const uint32_t want = nl < remaining ? nl : remaining;
                        cb = pool_used;
                        if (pool_used + want > L.pool_slots) overflow = 1;
                        pool_used += want;
                    }
                    if (n_nodes >= L.max_nodes) overflow = 1;
                    if (!overflow) {
                        const uint32_t child = n_nodes++;
                        store_state(states + child, lane, s);
                        if (LEARNED) store_rows4(frontier + (size_t)child * 28, lane, F);
                        if (bias_on) {
                            // _update_progressive_bias(parent, child): both values for the PARENT's player; the child's
                            // feature block also gives its own value, which the simulation below reuses
                            snapshot_features_block(s, F, (int)s.move_count, prm.leaf_max_turns, lane, s_pc, s_feat);
                            const double for_parent = win_probability_any(s_feat, prm.leaf_model, prm.leaf_model_kind, p, lane);
                            leaf_value = win_probability_any(s_feat, prm.leaf_model, prm.leaf_model_kind, s.to_move, lane);
                            have_value = true;
                            if (lane == 0) { node_prior[child] = __dsub_rn(for_parent, node_value[node]); node_value[child] = leaf_value; }
                            __syncwarp();
                        }
                        if (lane == 0) {
                            child_base[node] = cb;
                            pool[cb + ne] = child;
                            n_expanded[node] = (uint16_t)(ne + 1);
                            parent[child] = (int32_t)node; visits[child] = 0; reward[child] = 0;
                            child_base[child] = 0xFFFFFFFFu; n_legal[child] = (uint16_t)tc.n; n_expanded[child] = 0;
                            move_of[child] = (uint16_t)id;
                        }
                        node = child;
                        depth++;
                        __syncwarp();
                    }
                }
            }
            }   // lead && !overflow
            // ---- simulation (mcts_agent.py:408-437); `s` holds the state of `node`.  The transposition table is probed
            // by the tree's leading warp; a leaf-batched search caches the SUM of a leaf's W rewards.
            long long r = 0;
            double rd = 0.0;                         // the reward of learned evaluations
            bool cached = false;
            uint64_t key = 0;
            if (tt.keys && lead && !overflow) {
                key = zobrist_of(s, tt.zobrist, lane) ^ salt;
                double rv = 0.0;
                tt_probes++;
                if (tt_lookup(tt.keys, tt.rewards, tt.log2_slots, key, rv)) { r = (long long)rv; rd = rv; cached = true; tt_hits++; }
            }
            if (BATCH) {
                if (lead && lane == 0) { sh_node = node; sh_stop = overflow; sh_cached = cached ? 1u : 0u; }
                __syncthreads();                    // the leaf's state (global) and its index are visible to all warps
                overflow = sh_stop; node = sh_node; cached = sh_cached != 0u;
                if (!lead && !overflow && !cached) load_state(states + node, lane, s);
            }
            if (overflow) break;
            if (LEARNED) {
                if (!cached) {
                    if (prm.reward_mode == 4) {
                        // _rollout with potential-based shaping (mcts_agent.py:470-554): Phi of the leaf, the rollout
                        // (frontier sets follow every placement), and on a TRUNCATED rollout w * (gamma * Phi(end) - Phi(leaf))
                        const uint32_t me = s.to_move;
                        const int32_t before = board_score_of(s, me);
                        snapshot_features_block(s, F, (int)s.move_count, prm.leaf_max_turns, lane, s_pc, s_feat);
                        const double phi0 = potential_of(win_probability_any(s_feat, prm.leaf_model, prm.leaf_model_kind, me, lane), prm.potential_mode);
                        __syncwarp();
                        const int limit = prm.max_rollout_moves > 0 ? prm.max_rollout_moves : 1 << 30;
                        int moves_made = 0;
                        for (int ply = 0; ply < limit; ply++) {
                            const uint32_t p = s.to_move;
                            int o, row, col;
                            if (POLICY == BK_POLICY_HEURISTIC) {
                                uint32_t id = 0, idx = 0;
                                const uint32_t before_draws = draws;
                                const uint32_t rnd = next_draw();
                                if (heuristic_pick(s, p, rnd, hs, lane, false, id, idx) == 0) { draws = before_draws; break; }
                                o = (int)(id / 400u); row = (int)((id % 400u) / 20u); col = (int)(id % 20u);
                            } else {
                                uint32_t A, C; make_masks(s, p, lane, A, C);
                                const Totals tot = gen_totals<1>(A, C, sel4(s.used, p), lane);
                                if (tot.n == 0) break;
                                const uint32_t rnd = next_draw();
                                select_move(tot, __umulhi(rnd, tot.n), A, C, lane, o, row, col);
                            }
                            const uint32_t own_before = sel4(s.P, p);
                            place_move(s, p, o, row, col, lane);
                            update_frontier(s, p, sel4(s.P, p) & ~own_before, lane, F);
                            moves_made++;
                        }
                        rd = (double)(board_score_of(s, me) - before);
                        if (moves_made >= limit) {
                            snapshot_features_block(s, F, (int)s.move_count, prm.leaf_max_turns, lane, s_pc, s_feat);
                            const double phi1 = potential_of(win_probability_any(s_feat, prm.leaf_model, prm.leaf_model_kind, me, lane), prm.potential_mode);
                            __syncwarp();
                            rd = __dadd_rn(rd, __dmul_rn(prm.shaping_weight, __dsub_rn(__dmul_rn(prm.shaping_gamma, phi1), phi0)));
                        }
                    }
                    // _evaluate_leaf(node.board, node.player) -> predict_player_win_probability (learned_evaluator.py:79-143)
                    else if (have_value) rd = leaf_value;
                    else {
                        snapshot_features_block(s, F, (int)s.move_count, prm.leaf_max_turns, lane, s_pc, s_feat);
                        rd = win_probability_any(s_feat, prm.leaf_model, prm.leaf_model_kind, s.to_move, lane);
                        __syncwarp();
                    }
                    if (tt.keys && lane == 0) tt_insert(tt.keys, tt.rewards, tt.stamps, tt.log2_slots, tt.occupied, key, rd);
                }
            } else if (!cached && prm.reward_mode == 2) {
                // learned leaf evaluation in place of the rollout (mcts_agent.py:427-451) with the reference's dummy
                // artifact (learned_evaluator.py:51-55,129-131): every leaf is worth 0.5, counted here as one half
                r = 1;
                if (!BATCH && tt.keys && lane == 0) tt_insert(tt.keys, tt.rewards, tt.stamps, tt.log2_slots, tt.occupied, key, (double)r);
            } else if (!cached) {
                const uint32_t me = prm.reward_mode ? root_player : s.to_move;      // node.player (mcts_agent.py:483)
                const int32_t before = prm.reward_mode ? root_score : board_score_of(s, me);
                const int limit = prm.max_rollout_moves > 0 ? prm.max_rollout_moves : 1 << 30;
                for (int ply = 0; ply < limit; ply++) {
                    const uint32_t p = s.to_move;
                    if (POLICY == BK_POLICY_HEURISTIC) {                   // HeuristicAgent, the reference's default (:275-276)
                        uint32_t id = 0, idx = 0;
                        // a turn without a move consumes no draw: peek, then commit
                        const uint32_t before_draws = draws;
                        const uint32_t rnd = next_draw();
                        if (heuristic_pick(s, p, rnd, hs, lane, false, id, idx) == 0) { draws = before_draws; break; }
                        place_move(s, p, (int)(id / 400u), (int)((id % 400u) / 20u), (int)(id % 20u), lane);
                        continue;
                    }
                    uint32_t A, C; make_masks(s, p, lane, A, C);
                    const Totals tot = gen_totals<BATCH ? BK_MCTS_MAX_WARPS : 1>(A, C, sel4(s.used, p), lane);
                    if (tot.n == 0) break;                       // no passes inside a rollout (:488-489)
                    const uint32_t rnd = next_draw();
                    int o, row, col;
                    select_move(tot, __umulhi(rnd, tot.n), A, C, lane, o, row, col);
                    place_move(s, p, o, row, col, lane);
                }
                r = (long long)(board_score_of(s, me) - before);
                if (!BATCH && tt.keys && lane == 0) tt_insert(tt.keys, tt.rewards, tt.stamps, tt.log2_slots, tt.occupied, key, (double)r);
            }
            if (BATCH) {
                if (lane == 0) sh_r[wib] = r;       // (cached: the leading warp holds the cached sum, the others 0)
                __syncthreads();
                if (lead) {
                    r = 0; for (int w = 0; w < W; w++) r += sh_r[w];
                    if (tt.keys && !cached && lane == 0) tt_insert(tt.keys, tt.rewards, tt.stamps, tt.log2_slots, tt.occupied, key, (double)r);
                }
            }
            // ---- back-propagation (mcts_agent.py:572-582): W simulations at once in batch mode
            if (lead && lane == 0) {
                if (LEARNED) { for (int32_t v = (int32_t)node; v >= 0; v = parent[v]) { visits[v] += 1; reward_d[v] = __dadd_rn(reward_d[v], rd); } }
                else { for (int32_t v = (int32_t)node; v >= 0; v = parent[v]) { visits[v] += W; reward[v] += r; } }
            }
            if (lead) iter_done++;
            __syncwarp();
        }

        if (!lead && lane == 0) hdr->draws_x[wib - 1] = draws;
        if (lead && lane == 0) {
            hdr->n_nodes = n_nodes; hdr->pool_used = pool_used; hdr->iterations_done = iter_done;
            hdr->draws = draws; hdr->overflow = overflow; hdr->root_children = n_expanded[0];
            hdr->root_player = root_player; hdr->root_score = root_score;
            if (out_status) out_status[tree] = overflow;
            if (out_counts) out_counts[tree] = n_legal[0];
            if (tt.keys && tt_probes) {           // TranspositionTable.access_count / hit_count (mcts/zobrist.py:181-185)
                atomicAdd(tt.occupied + 2, (unsigned long long)tt_hits);
                atomicAdd(tt.occupied + 3, (unsigned long long)tt_probes);
            }
        }
        __syncwarp();
        // ---- root statistics by ascending legal-move rank (child k of the root is rank n_legal-1-k)
        if (lead) {
            const uint32_t nl0 = n_legal[0], ne0 = n_expanded[0], cb0 = child_base[0];
            for (uint32_t rank = lane; rank < (uint32_t)max_root_moves; rank += 32) {
                int32_t v = 0; long long w = 0; uint16_t m = 0xFFFF;
                if (rank < nl0) {
                    const uint32_t k = nl0 - 1 - rank;
                    if (k < ne0) { const uint32_t ch = pool[cb0 + k]; v = visits[ch]; w = reward[ch]; m = move_of[ch]; }
                }
                const size_t oi = (size_t)tree * max_root_moves + rank;
                if (out_visits) out_visits[oi] = v;
                if (out_reward) out_reward[oi] = w;
                if (out_moves) out_moves[oi] = m;
            }
        }
        if (BATCH) __syncthreads();
    }
}
