// mcts_wide.cuh -- K6: ONE search tree spread over the whole GPU ("round-synchronous" tree-parallel UCT).
//
// k_mcts (mcts.cuh) runs the reference's sequential iteration, one warp (or one CTA) per tree: exact, but a single move of
// 1,000 simulations keeps 1 of 148 SMs busy.  The callers that ask for ONE move at a time (FastMCTSAgent.think, POST /think,
// the arena's gameplay agents) get this search instead.  It is a different -- deterministic, oracle-restated -- schedule
// of the same ingredients (UCB1 selection, pop-last expansion in ascending move order, Philox rollouts, integer rewards):
//
//   round:  PLAN   one warp walks the tree `width` times.  A walk descends through fully expanded nodes by
//                  u = W * (1/m) + (c * sqrt(L)) * (1/sqrt(n)),   W, m = reward sum and number of the child's FINISHED
//                  simulations (0 while it has none), n = m + its pending ones ("watch the unobserved": pending
//                  simulations shrink the exploration term, they do not count as losses),
//                  L = ln(N) (or 2 ln N, fast_mcts_agent.py:52) with N = finished + pending of the parent; first maximum wins.
//                  At a node with an untried move it claims the last one (mcts_agent.py:124) for a NEW child; at a terminal /
//                  depth-limit node it schedules another simulation of that node.  A child created in this round is a LEAF
//                  until the round ends (its state does not exist yet): a walk that reaches it schedules one more
//                  simulation of it, which re-derives the child from (parent, rank) like the creating one.
//           WORK   every scheduled simulation runs on its own warp anywhere on the GPU: build the new child's state
//                  (re-derive the claimed move by rank, place it, count its legal moves), roll out with the Philox stream
//                  of (seed, tree, simulation number), add visit and reward to the leaf and all its ancestors (integer
//                  atomics: order-independent, hence deterministic).
//           DEALT ROUNDS  walks that need no selection are not walked: while the root has untried moves every walk claims the
//                  next one, so they are written down in one sweep; and a root that is fully expanded hands the rest of the
//                  round (flat bandit: always; full tree: from BK_WIDE_DEAL_MIN walks) to its children at once -- the W largest
//                  values of the children's non-increasing sequences u_k(j) with the exploration scale frozen at the round's
//                  final count, i.e. the greedy "next walk to the first maximal child", found by a threshold search.  A flat
//                  bandit is then done; in a full tree walk i goes on below its child, child after child, and runs of
//                  expansions at one child are again written down at once.
//   The result depends on (seed, tree id, width, the sequence of search calls) only -- not on the grid or the SM count.
//
// Restated sequentially in oracle/blokus_oracle.c (orc_mcts_wide); tests/test_gpu_mcts_wide.py compares bit for bit.
#pragma once
#include "mcts.cuh"

#define BK_WIDE_NONE 0xFFFFFFFFu
#define BK_WIDE_STREAM 0x57u        /* Philox stream id of the wide search's rollouts */
#define BK_WIDE_MAX_WIDTH 4096
#define BK_WIDE_DEAL_MIN 128u       /* full trees: walks left in the round from which the root deals them out at once */
#define BK_WIDE_PENDING 0xFFFFu     /* n_legal of a node whose state the current round is still building */

struct WideLayout {
    uint32_t max_nodes, pool_slots, width, table_len;
    size_t off_states, off_reward, off_q, off_inv, off_parent, off_visits, off_virt, off_child_base, off_pool,
           off_n_legal, off_n_expanded, off_move, off_item_leaf, off_item_parent, off_item_rank, off_header,
           off_rcp, off_invtab, off_stab, total;
};

__host__ __device__ inline WideLayout wide_layout(int sims_total, int width, int max_depth) {
    WideLayout L;
    const size_t T = (size_t)(sims_total > 0 ? sims_total : 1);          // simulations (sizes the tables)
    const size_t TN = (max_depth == 1 && T > 2048) ? 2048 : T;           // nodes: a flat bandit has at most 2048 + root
    L.max_nodes = (uint32_t)(TN + 1);
    size_t pool = TN * (TN + 1) / 2 + 1;
    if (pool > TN * 2048) pool = TN * 2048;
    if (max_depth == 1 && pool > 2049) pool = 2049;                      // only the root has children
    L.pool_slots = (uint32_t)(pool > 0xFFFFFF00ull ? 0xFFFFFF00ull : pool);
    L.width = (uint32_t)(width < 1 ? 1 : (width > BK_WIDE_MAX_WIDTH ? BK_WIDE_MAX_WIDTH : width));
    size_t o = 0;
    L.off_states = o;       o += (size_t)L.max_nodes * sizeof(bk_state);
    L.off_reward = o;       o += (size_t)L.max_nodes * 8;
    L.table_len = (uint32_t)(T + L.width + 2);              // simulation counts 0 .. T + width (finished + pending)
    L.off_rcp = o;          o += (size_t)L.table_len * 8;
    L.off_q = o;            o += (size_t)L.max_nodes * 4;
    L.off_inv = o;          o += (size_t)L.max_nodes * 4;
    L.off_invtab = o;       o += (size_t)L.table_len * 4;
    L.off_stab = o;         o += (size_t)L.table_len * 4;
    L.off_parent = o;       o += (size_t)L.max_nodes * 4;
    L.off_visits = o;       o += (size_t)L.max_nodes * 4;
    L.off_virt = o;         o += (size_t)L.max_nodes * 4;
    L.off_child_base = o;   o += (size_t)L.max_nodes * 4;
    L.off_pool = o;         o += (size_t)L.pool_slots * 4;
    L.off_item_leaf = o;    o += (size_t)L.width * 4;
    L.off_item_parent = o;  o += (size_t)L.width * 4;
    L.off_n_legal = o;      o += (size_t)L.max_nodes * 2;
    L.off_n_expanded = o;   o += (size_t)L.max_nodes * 2;
    L.off_move = o;         o += (size_t)L.max_nodes * 2;
    L.off_item_rank = o;    o += (size_t)L.width * 2;
    o = bk_align_up(o, 8);
    L.off_header = o;       o += 128;
    L.total = bk_align_up(o, 256);
    return L;
}

struct WideHeader {            // persistent between the launches of one search and between resumed calls
    uint32_t n_nodes, pool_used, sims_done, round, round_items, ticket, sims_base, overflow, root_player;
    int32_t root_score;
    uint32_t inited;
};
static_assert(sizeof(WideHeader) <= 128, "WideHeader must fit its workspace slot");

struct WideView {              // typed pointers into one tree's workspace slice
    bk_state* states; long long* reward; float* q; float* inv; int32_t* parent; int32_t* visits; int32_t* virt;
    uint32_t* child_base; uint32_t* pool; uint16_t* n_legal; uint16_t* n_expanded; uint16_t* move_of;
    uint32_t* item_leaf; uint32_t* item_parent; uint16_t* item_rank; WideHeader* hdr;
    const double* rcp; const float* invtab; const float* stab;
};
__device__ __forceinline__ WideView wide_view(unsigned char* base, const WideLayout& L) {
    WideView v;
    v.states = reinterpret_cast<bk_state*>(base + L.off_states);
    v.reward = reinterpret_cast<long long*>(base + L.off_reward);
    v.q = reinterpret_cast<float*>(base + L.off_q);
    v.inv = reinterpret_cast<float*>(base + L.off_inv);
    v.parent = reinterpret_cast<int32_t*>(base + L.off_parent);
    v.visits = reinterpret_cast<int32_t*>(base + L.off_visits);
    v.virt = reinterpret_cast<int32_t*>(base + L.off_virt);
    v.child_base = reinterpret_cast<uint32_t*>(base + L.off_child_base);
    v.pool = reinterpret_cast<uint32_t*>(base + L.off_pool);
    v.n_legal = reinterpret_cast<uint16_t*>(base + L.off_n_legal);
    v.n_expanded = reinterpret_cast<uint16_t*>(base + L.off_n_expanded);
    v.move_of = reinterpret_cast<uint16_t*>(base + L.off_move);
    v.item_leaf = reinterpret_cast<uint32_t*>(base + L.off_item_leaf);
    v.item_parent = reinterpret_cast<uint32_t*>(base + L.off_item_parent);
    v.item_rank = reinterpret_cast<uint16_t*>(base + L.off_item_rank);
    v.hdr = reinterpret_cast<WideHeader*>(base + L.off_header);
    v.rcp = reinterpret_cast<const double*>(base + L.off_rcp);
    v.invtab = reinterpret_cast<const float*>(base + L.off_invtab);
    v.stab = reinterpret_cast<const float*>(base + L.off_stab);
    return v;
}

// Selection terms.  Everything expensive is tabulated once per search (k_wide_tables) in float64 and rounded to float32:
// for m finished simulations of a child with reward sum w, n = m + its pending ones and N of its parent
//     u = fma(s[N], inv[n], q)   in float32,
//     q = float((w * unit) * rcp[m]),  rcp[m] = 1/m,  inv[n] = float(1/sqrt(n)),  s[N] = float(c * sqrt(L(N)))
// so that a walk is loads, one fused multiply-add per child and one warp arg-max.  (Single precision on purpose: the walks
// are one long dependent chain on a single warp, and float64 arithmetic has several times the latency; q is a mean of
// integer rewards below 2^7, 24 bits resolve differences far below the exploration term.)
__global__ void __launch_bounds__(256)
k_wide_tables(int n_trees, bk_mcts_params prm, unsigned char* __restrict__ workspace) {
    const int tree = blockIdx.y;
    if (tree >= n_trees) return;
    const WideLayout L = wide_layout(prm.iterations_total, prm.round_width, prm.max_depth);
    unsigned char* base = workspace + (size_t)tree * L.total;
    double* rcp = reinterpret_cast<double*>(base + L.off_rcp);
    float* inv = reinterpret_cast<float*>(base + L.off_invtab);
    float* stab = reinterpret_cast<float*>(base + L.off_stab);
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < L.table_len; n += gridDim.x * blockDim.x) {
        const double dn = (double)(n ? n : 1u);
        rcp[n] = __ddiv_rn(1.0, dn);
        inv[n] = __double2float_rn(__ddiv_rn(1.0, sqrt(dn)));
        stab[n] = __double2float_rn(__dmul_rn(prm.exploration, sqrt(__dmul_rn(log(dn), prm.ucb_variant ? 2.0 : 1.0))));
    }
}

// 32-bit key whose unsigned order is the order of the floats (finite values and infinities; never 0)
__device__ __forceinline__ uint32_t wide_key(float u) {
    const uint32_t b = __float_as_uint(u);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float wide_unkey(uint32_t k) {
    return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k);
}

// ---- PLAN: one CTA per tree ----------------------------------------------------------------------------------------------
// The walks of a round are one dependent chain (each walk sees the pending counts of the walks before it), so the time of
// a round is (number of walks) x (latency of one walk).  Two things keep that latency short:
//   * a shared-memory image of the first BK_WIDE_NC nodes -- the root, all its children (created first: nodes 1..n) and the
//     earliest grandchildren; a flat bandit never has more -- loaded at the start of the round; pending counts and
//     exploration terms of these nodes are written back in one sweep at the end; later nodes are read from global memory;
//   * BK_WIDE_PLAN_THREADS threads score the children of a node together (one child per thread up to 256 children; a node
//     with few children is scored by every warp redundantly, which needs no exchange), two warp reductions + one barrier
//     + two warp reductions give the first maximum; everything after the choice (claiming the move, the item, the pending
//     counts along the path) is warp 0's work, and one barrier ends the walk.
#define BK_WIDE_NC 2049
// Measured on B200 (profiles/r02_wide_latency.md): one warp 0.65 us per walk; eight warps sharing the scoring 0.8 us -- the
// two barriers per walk cost more than the shorter scoring loop saves, the walk is latency-bound outside the loop.  So the
// planner is ONE warp (the cooperative path stays compiled for wider CTAs); callers that want more than one planner run
// several root-parallel trees side by side (n_trees > 1: one planner CTA each, the work grid shared).
#define BK_WIDE_PLAN_THREADS 32
#define BK_WIDE_PLAN_WARPS (BK_WIDE_PLAN_THREADS / 32)
#define BK_WIDE_SOLO (BK_WIDE_PLAN_THREADS == 32 ? 0xFFFFFFFFu : 64u)   /* children up to which every warp scores all of them */
struct WideSmem {
    float q[BK_WIDE_NC], inv[BK_WIDE_NC];
    int32_t vis[BK_WIDE_NC], virt[BK_WIDE_NC], par[BK_WIDE_NC];
    uint32_t cb[BK_WIDE_NC];
    uint16_t nl[BK_WIDE_NC], ne[BK_WIDE_NC], mv[BK_WIDE_NC];
    uint32_t cand_key[2][BK_WIDE_PLAN_WARPS], cand_k[2][BK_WIDE_PLAN_WARPS];
    uint16_t give[2048], first[2048];       // flat batch rounds: walks given to root child k, position of its first item
    uint32_t n_nodes, pool_used, overflow, stop;
};
#define BK_WIDE_PLAN_SMEM ((int)sizeof(WideSmem))

__global__ void __launch_bounds__(BK_WIDE_PLAN_THREADS)
k_wide_plan(const bk_state* __restrict__ roots, int n_roots, int n_trees, bk_mcts_params prm, unsigned char* __restrict__ workspace,
            int fresh, uint32_t sims_target) {
    extern __shared__ __align__(16) unsigned char plan_smem[];
    WideSmem& S = *reinterpret_cast<WideSmem*>(plan_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tree = blockIdx.x;
    if (tree >= n_trees) return;
    const WideLayout L = wide_layout(prm.iterations_total, prm.round_width, prm.max_depth);
    const WideView V = wide_view(workspace + (size_t)tree * L.total, L);
    WideHeader* hdr = V.hdr;
    const double unit = prm.reward_mode == 2 ? 0.5 : 1.0;

    if (fresh) {
        if (warp == 0) {
            WarpState s;
            load_state(roots + (tree % n_roots), lane, s);
            uint32_t A, C; make_masks(s, s.to_move, lane, A, C);
            const Totals t0 = gen_totals<BK_MCTS_MAX_WARPS>(A, C, sel4(s.used, s.to_move), lane);
            store_state(V.states + 0, lane, s);
            const int32_t rs = board_score_of(s, s.to_move);
            if (lane == 0) {
                V.parent[0] = -1; V.visits[0] = 0; V.virt[0] = 0; V.reward[0] = 0; V.child_base[0] = BK_WIDE_NONE;
                V.n_legal[0] = (uint16_t)t0.n; V.n_expanded[0] = 0; V.move_of[0] = 0xFFFF; V.q[0] = 0.0f; V.inv[0] = 0.0f;
                hdr->n_nodes = 1; hdr->pool_used = 0; hdr->sims_done = 0; hdr->round = 0; hdr->round_items = 0; hdr->ticket = 0;
                hdr->sims_base = 0; hdr->overflow = 0; hdr->root_player = s.to_move; hdr->root_score = rs; hdr->inited = 1;
            }
        }
        __threadfence_block();
        __syncthreads();
    }
    uint32_t n_nodes = hdr->n_nodes, pool_used = hdr->pool_used, overflow = hdr->overflow;
    // ---- the previous round is complete: drop its pending counts and refresh the selection terms along its paths
    const uint32_t prev = hdr->round_items;
    for (uint32_t i = tid; i < prev; i += BK_WIDE_PLAN_THREADS) {
        for (int32_t v = (int32_t)V.item_leaf[i]; v > 0; v = V.parent[v]) {      // the root is nobody's child: no terms
            const int32_t n = V.visits[v];
            V.virt[v] = 0;
            V.q[v] = __double2float_rn(__dmul_rn(__dmul_rn((double)V.reward[v], unit), V.rcp[n]));
            V.inv[v] = V.invtab[n];
        }
    }
    __threadfence_block();
    __syncthreads();
    // ---- shared-memory image of the first nodes
    const uint32_t nc = n_nodes < BK_WIDE_NC ? n_nodes : BK_WIDE_NC;
    for (uint32_t v = tid; v < nc; v += BK_WIDE_PLAN_THREADS) {
        S.q[v] = V.q[v]; S.inv[v] = V.inv[v]; S.vis[v] = V.visits[v]; S.virt[v] = 0; S.par[v] = V.parent[v];
        S.cb[v] = V.child_base[v]; S.nl[v] = V.n_legal[v]; S.ne[v] = V.n_expanded[v]; S.mv[v] = V.move_of[v];
    }
    if (tid == 0) { S.n_nodes = n_nodes; S.pool_used = pool_used; S.overflow = overflow; S.stop = 0; }
    __syncthreads();
    const uint32_t sims_done = hdr->sims_done;
    const uint32_t room = sims_target > sims_done ? sims_target - sims_done : 0u;
    uint32_t want_items = overflow ? 0u : (room < L.width ? room : L.width);
    uint32_t n_items = 0, buf = 0;
    // ---- while the root has untried moves every walk stops there and claims the next one (no selection happens), so these
    // walks are written down at once: children n_nodes, n_nodes + 1, ... with ranks nl-1-ne, nl-2-ne, ... -- exactly what the
    // walks below would do one after the other (0.6 us each; a 237-move root costs 150 us that way).
    static_assert(BK_WIDE_PLAN_WARPS == 1, "the batched root expansions are written for one planner warp");
    if (n_items < want_items && !overflow && S.nl[0] != BK_WIDE_PENDING && S.ne[0] < S.nl[0]) {
        if (warp == 0) {
            const uint32_t nl = S.nl[0], ne = S.ne[0];
            uint32_t cb = S.cb[0];
            if (cb == BK_WIDE_NONE) {
                const uint32_t remaining = (uint32_t)prm.iterations_total - (sims_done + n_items);
                const uint32_t want = nl < remaining ? nl : remaining;
                cb = pool_used;
                if (pool_used + want > L.pool_slots) overflow = 1;
                pool_used += want;
            }
            uint32_t e = min(want_items - n_items, nl - ne);
            bool stop = false;
            if (overflow) { e = 0; stop = true; }
            else if (n_nodes + e > L.max_nodes) { e = L.max_nodes - n_nodes; overflow = 1; stop = true; }   // the walk after the last fitting one
            const float inv1 = __ldg(V.invtab + 1);
            __syncwarp();
            for (uint32_t i = lane; i < e; i += 32u) {
                const uint32_t child = n_nodes + i, k = ne + i, rank = nl - 1u - k;
                V.pool[cb + k] = child;
                V.parent[child] = 0; V.visits[child] = 0; V.virt[child] = 0; V.reward[child] = 0;
                V.child_base[child] = BK_WIDE_NONE; V.n_legal[child] = BK_WIDE_PENDING; V.n_expanded[child] = 0;
                V.move_of[child] = (uint16_t)rank; V.q[child] = 0.0f; V.inv[child] = 0.0f;
                if (child < BK_WIDE_NC) {
                    S.par[child] = 0; S.vis[child] = 0; S.virt[child] = 1; S.cb[child] = BK_WIDE_NONE;
                    S.nl[child] = BK_WIDE_PENDING; S.ne[child] = 0; S.mv[child] = (uint16_t)rank; S.q[child] = 0.0f; S.inv[child] = inv1;
                } else { V.virt[child] = 1; V.inv[child] = inv1; }
                V.item_leaf[n_items + i] = child; V.item_parent[n_items + i] = 0u; V.item_rank[n_items + i] = (uint16_t)rank;
            }
            __syncwarp();
            if (lane == 0) {
                if (e > 0) { V.child_base[0] = cb; S.cb[0] = cb; }
                V.n_expanded[0] = (uint16_t)(ne + e); S.ne[0] = (uint16_t)(ne + e);
                S.virt[0] += (int32_t)e;
                S.n_nodes = n_nodes + e; S.pool_used = pool_used; S.overflow = overflow; S.stop = stop ? 1u : 0u;
            }
            n_items += e;
        }
        __threadfence_block();
        __syncthreads();
        n_nodes = S.n_nodes; pool_used = S.pool_used; overflow = S.overflow;
        if (S.stop) want_items = n_items;
    }
    // ---- a root that is fully expanded -- at the start of the round, or once this round's expansions are written down --
    // deals the rest of the round's walks out to its children at once (flat bandit: that is the whole walk).  With the exploration scale frozen at the round's final
    // count, s = stab[finished + pending + W], the value of giving child k one more walk when it already has j of them is
    // u_k(j) = fma(s, invtab[m_k + p_k + j], q_k) (m finished, p pending simulations; a child created in this round has m = 0,
    // p = 1, q = 0), non-increasing in j, and the sequential rule "the next walk goes to the first maximal child" picks
    // exactly the W largest values (ties: lower k first, then lower j).  That
    // set is found without walking: a bit-by-bit search for the largest key T with #{(k, j): key(u_k(j)) >= T} >= W, every
    // child's count by a binary search over j; children take everything above T and the values equal to T are shared out in
    // child order.  Cost O(children * log W * 32) per round instead of O(W * children).  The items are numbered child by child.
    const bool flat = prm.max_depth == 1;
    bool dealt = false;                                              // full tree: walk i of the round starts at root child item_leaf[i]
    uint32_t dealt0 = 0;                                             // position of the first dealt walk
    // (the dealing costs 60-100 us whatever the width and saves most of a walk's 0.65-0.85 us:
    //  full trees use it for rounds of at least BK_WIDE_DEAL_MIN walks -- part of the schedule, the oracle has the same rule)
    if (n_items < want_items && (flat || want_items - n_items >= BK_WIDE_DEAL_MIN) && S.nl[0] != BK_WIDE_PENDING && S.nl[0] > 0 &&
        S.ne[0] == S.nl[0] && S.nl[0] <= 2048) {
        if (warp == 0) {
            const uint32_t ne0 = S.ne[0], W = want_items - n_items, item0 = n_items;
            const float s = __ldg(V.stab + S.vis[0] + S.virt[0] + (int32_t)W);
            // walks of child k whose key is >= tau: first j with key(u_k(j)) < tau (keys do not increase with j).  The
            // boundary is GUESSED in closed form -- fma(s, 1/sqrt(m + j), q) >= t  <=>  m + j <= (s / (t - q))^2 up to rounding --
            // and then settled with the real table values (the answer is the exact boundary of the monotone predicate, the
            // same a binary search finds), which takes two or three dependent loads instead of log2(W).
            const float inv_s = s > 0.0f ? 1.0f / s : 0.0f;
            auto guess = [&](float q, int32_t m, float t) -> uint32_t {
                const float x = (t - q) * inv_s;                           // needs 1/sqrt(m + j) >= x
                if (!(x > 0.0f) || !(s > 0.0f)) return W;                  // every j qualifies (or s == 0: settled below)
                const float r = 1.0f / (x * x) - (float)m + 1.0f;
                return r <= 0.0f ? 0u : (r >= (float)W ? W : (uint32_t)r);
            };
            // one step towards the boundary in [lo, hi] (ok(j) for j < lo, !ok(j) for j >= hi): test the guess or its left neighbour
            auto settle_step = [&](const float* it, float q, uint32_t tau, uint32_t& lo, uint32_t& hi, uint32_t& g, int step) {
                if (step >= 6) g = (lo + hi) >> 1;                         // a bad guess: plain bisection from here
                g = g < lo ? lo : (g > hi ? hi : g);
                const uint32_t j = g == hi ? g - 1u : g;                   // g == hi > lo: is the left neighbour still in?
                const bool in = wide_key(__fmaf_rn(s, __ldg(it + j), q)) >= tau;
                if (g == hi) { if (in) lo = hi; else { hi = g - 1u; g = hi; } }
                else if (in) { lo = g + 1u; g = lo; }
                else { hi = g; if (g > lo) g = g - 1u; }
            };
            auto count_ge = [&](uint32_t k, uint32_t tau) -> uint32_t {
                const float q = S.q[k + 1];
                const int32_t m = S.vis[k + 1] + S.virt[k + 1];
                uint32_t lo = 0, hi = W, g = guess(q, m, wide_unkey(tau));
                for (int step = 0; lo < hi; step++) settle_step(V.invtab + m, q, tau, lo, hi, g, step);
                return lo;
            };
            // the same for all children, eight per lane at a time: eight independent chains keep eight table loads in flight
            auto total_ge = [&](uint32_t tau) -> uint32_t {
                const float t = wide_unkey(tau);
                uint32_t sum = 0;
                for (uint32_t base = lane; base < ne0; base += 256u) {
                    uint32_t lo[8], hi[8], g[8]; float q[8]; const float* it[8];
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        const uint32_t k = base + 32u * (uint32_t)r;
                        const bool valid = k < ne0;
                        const int32_t m = valid ? S.vis[k + 1] + S.virt[k + 1] : 1;
                        q[r] = valid ? S.q[k + 1] : 0.0f;
                        it[r] = V.invtab + m;
                        lo[r] = 0u; hi[r] = valid ? W : 0u; g[r] = guess(q[r], m, t);
                    }
                    bool more = true;
                    for (int step = 0; more; step++) {
                        more = false;
#pragma unroll
                        for (int r = 0; r < 8; r++)
                            if (lo[r] < hi[r]) { settle_step(it[r], q[r], tau, lo[r], hi[r], g[r], step); more = true; }
                    }
#pragma unroll
                    for (int r = 0; r < 8; r++) sum += lo[r];
                }
                return __reduce_add_sync(BK_FULL, sum);
            };
            // the threshold lies between the smallest last value and the largest first value of the children
            uint32_t kmax = 0, kmin = 0xFFFFFFFFu;
            for (uint32_t k = lane; k < ne0; k += 32) {
                const float q = S.q[k + 1];
                const float* it = V.invtab + S.vis[k + 1] + S.virt[k + 1];
                kmax = max(kmax, wide_key(__fmaf_rn(s, __ldg(it), q)));
                kmin = min(kmin, wide_key(__fmaf_rn(s, __ldg(it + (W - 1u)), q)));
            }
            kmax = __reduce_max_sync(BK_FULL, kmax); kmin = __reduce_min_sync(BK_FULL, kmin);
            uint32_t tau = kmin, top = kmax;                                             // total_ge(kmin) = children * W >= W
            bool exact = false;                                                          // a key with exactly W values at or above it
            while (tau < top) {
                const uint32_t mid = tau + ((top - tau + 1u) >> 1);
                const uint32_t c = total_ge(mid);
                if (c == W) { tau = mid - 1u; exact = true; break; }                     // those W values ARE the W largest: done
                if (c > W) tau = mid; else top = mid - 1u;
            }
            const uint32_t above = exact ? W : (tau == 0xFFFFFFFFu ? 0u : total_ge(tau + 1u));   // walks with a key strictly above tau
            uint32_t left = W - above, at = 0;                                           // still to share out; next item position
            for (uint32_t row = 0; row * 32u < ne0; row++) {
                const uint32_t k = row * 32u + (uint32_t)lane;
                uint32_t hi_cnt = 0, eq = 0;
                if (k < ne0) {
                    hi_cnt = tau == 0xFFFFFFFFu ? 0u : count_ge(k, tau + 1u);
                    eq = count_ge(k, tau) - hi_cnt;
                }
                uint32_t incl = eq;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(BK_FULL, incl, d); if (lane >= d) incl += v; }
                const uint32_t before = incl - eq;                                       // equal-valued walks of lower children in this row
                const uint32_t take = before >= left ? 0u : (eq < left - before ? eq : left - before);
                const uint32_t mine = hi_cnt + take;
                const uint32_t row_eq = __shfl_sync(BK_FULL, incl, 31);
                left -= row_eq < left ? row_eq : left;
                uint32_t pincl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(BK_FULL, pincl, d); if (lane >= d) pincl += v; }
                if (k < ne0) {
                    S.give[k] = (uint16_t)mine; S.first[k] = (uint16_t)(at + pincl - mine);
                    if (flat && mine) { S.virt[k + 1] += (int32_t)mine; S.inv[k + 1] = __ldg(V.invtab + S.vis[k + 1] + S.virt[k + 1]); }
                }
                at += __shfl_sync(BK_FULL, pincl, 31);
            }
            __syncwarp();
            // items: position t belongs to the last child whose first position is <= t
            for (uint32_t t = lane; t < W; t += 32) {
                uint32_t lo = 0, hi = ne0 - 1;
                while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (S.first[mid] <= t) lo = mid; else hi = mid - 1; }
                const bool pend = S.nl[lo + 1] == BK_WIDE_PENDING;       // created in this round: the work kernel re-derives it
                V.item_leaf[item0 + t] = lo + 1; V.item_parent[item0 + t] = pend ? 0u : BK_WIDE_NONE;
                V.item_rank[item0 + t] = pend ? S.mv[lo + 1] : (uint16_t)0;
            }
            if (flat && lane == 0) S.virt[0] += (int32_t)W;
        }
        // The flat bandit is done with the round.  In a full tree the dealing replaces only the ROOT's selection: walk i goes
        // on below the child it was dealt to (item_leaf[i] holds that child until the walk writes its own leaf there), one
        // after the other as before -- the scan of all root children per walk is what is saved.
        if (flat) n_items = want_items; else { dealt = true; dealt0 = n_items; }
        __threadfence_block();
        __syncthreads();
    }
    for (; n_items < want_items; n_items++) {
        // ---- a dealt round: the walks of one root child are consecutive.  While that child has untried moves each of them
        // stops there and claims the next one; while it is itself being built each of them is one more simulation of it.
        // Either way the walks need no selection and are written down at once (what they would do one after the other).
        if (dealt) {
            const uint32_t nd = V.item_leaf[n_items], k = nd - 1u;            // a root child: node k + 1 < BK_WIDE_NC
            if (n_items == dealt0 + S.first[k] && warp == 0) {
                const uint32_t g = S.give[k], nl = S.nl[nd], ne = S.ne[nd];
                uint32_t e = 0;
                bool stop = false;
                if (nl == BK_WIDE_PENDING) {
                    e = g;
                    const uint16_t rank = S.mv[nd];
                    for (uint32_t i = lane; i < e; i += 32u) {
                        V.item_leaf[n_items + i] = nd; V.item_parent[n_items + i] = 0u; V.item_rank[n_items + i] = rank;
                    }
                } else if (ne < nl && (prm.max_depth <= 0 || 1 < prm.max_depth)) {
                    uint32_t cb = S.cb[nd];
                    if (cb == BK_WIDE_NONE) {
                        const uint32_t remaining = (uint32_t)prm.iterations_total - (sims_done + n_items);
                        const uint32_t want = nl < remaining ? nl : remaining;
                        cb = pool_used;
                        if (pool_used + want > L.pool_slots) overflow = 1;
                        pool_used += want;
                    }
                    e = min(g, nl - ne);
                    if (overflow) { e = 0; stop = true; }
                    else if (n_nodes + e > L.max_nodes) { e = L.max_nodes - n_nodes; overflow = 1; stop = true; }
                    const float inv1 = __ldg(V.invtab + 1);
                    __syncwarp();
                    for (uint32_t i = lane; i < e; i += 32u) {
                        const uint32_t child = n_nodes + i, kk = ne + i, rank = nl - 1u - kk;
                        V.pool[cb + kk] = child;
                        V.parent[child] = (int32_t)nd; V.visits[child] = 0; V.reward[child] = 0;
                        V.child_base[child] = BK_WIDE_NONE; V.n_legal[child] = BK_WIDE_PENDING; V.n_expanded[child] = 0;
                        V.move_of[child] = (uint16_t)rank; V.q[child] = 0.0f;
                        if (child < BK_WIDE_NC) {
                            V.virt[child] = 0; V.inv[child] = 0.0f;
                            S.par[child] = (int32_t)nd; S.vis[child] = 0; S.virt[child] = 1; S.cb[child] = BK_WIDE_NONE;
                            S.nl[child] = BK_WIDE_PENDING; S.ne[child] = 0; S.mv[child] = (uint16_t)rank; S.q[child] = 0.0f; S.inv[child] = inv1;
                        } else { V.virt[child] = 1; V.inv[child] = inv1; }
                        V.item_leaf[n_items + i] = child; V.item_parent[n_items + i] = nd; V.item_rank[n_items + i] = (uint16_t)rank;
                    }
                    __syncwarp();
                    if (lane == 0 && e > 0) { V.child_base[nd] = cb; S.cb[nd] = cb; V.n_expanded[nd] = (uint16_t)(ne + e); S.ne[nd] = (uint16_t)(ne + e); }
                    n_nodes += e;
                }
                if (e > 0 || stop) {
                    __syncwarp();
                    if (lane == 0) {
                        S.virt[nd] += (int32_t)e; S.inv[nd] = __ldg(V.invtab + S.vis[nd] + S.virt[nd]); S.virt[0] += (int32_t)e;
                        S.n_nodes = n_nodes; S.pool_used = pool_used; S.overflow = overflow;
                    }
                    __syncwarp();
                    if (stop) { n_items += e; break; }
                    if (e > 0) { n_items += e - 1u; continue; }              // (the loop adds the last one)
                }
            }
        }
        // ---- descent: all threads, uniform control flow
        uint32_t node = dealt ? V.item_leaf[n_items] : 0u, depth = dealt ? 1u : 0u;
        bool may_expand = false, pending = false;
        for (;;) {
            const bool in = node < BK_WIDE_NC;
            const uint32_t nl = in ? S.nl[node] : V.n_legal[node], ne = in ? S.ne[node] : V.n_expanded[node];
            if (nl == BK_WIDE_PENDING) { pending = true; break; }        // created in this round: a leaf until the round ends
            may_expand = ne < nl && (prm.max_depth <= 0 || (int)depth < prm.max_depth);
            if (may_expand || ne == 0) break;
            const int32_t big_n = in ? S.vis[node] + S.virt[node] : V.visits[node] + V.virt[node];
            const float s = __ldg(V.stab + big_n);
            const uint32_t cb = in ? S.cb[node] : V.child_base[node];
            const bool solo = ne <= BK_WIDE_SOLO;                        // every warp scores all children: no exchange
            const uint32_t first = solo ? (uint32_t)lane : (uint32_t)tid, step = solo ? 32u : (uint32_t)BK_WIDE_PLAN_THREADS;
            float best = 0.0f; uint32_t best_k = BK_WIDE_NONE;
            if (node == 0) {
                // the root's children are nodes 1 .. ne in creation order (nothing else is created while it has untried moves)
                for (uint32_t k = first; k < ne; k += step) {
                    const float u = __fmaf_rn(s, S.inv[k + 1], S.q[k + 1]);
                    if (best_k == BK_WIDE_NONE || u > best) { best = u; best_k = k; }
                }
            } else {
                for (uint32_t k = first; k < ne; k += step) {
                    const uint32_t ch = V.pool[cb + k];
                    const bool cin = ch < BK_WIDE_NC;
                    const float u = __fmaf_rn(s, cin ? S.inv[ch] : V.inv[ch], cin ? S.q[ch] : V.q[ch]);
                    if (best_k == BK_WIDE_NONE || u > best) { best = u; best_k = k; }
                }
            }
            // first maximum: largest key, then smallest k
            uint32_t key = best_k == BK_WIDE_NONE ? 0u : wide_key(best);
            uint32_t mk = __reduce_max_sync(BK_FULL, key);
            uint32_t kk = __reduce_min_sync(BK_FULL, key == mk ? best_k : BK_WIDE_NONE);
            if (!solo) {
                if (lane == 0) { S.cand_key[buf][warp] = mk; S.cand_k[buf][warp] = kk; }
                __syncthreads();
                key = lane < BK_WIDE_PLAN_WARPS ? S.cand_key[buf][lane] : 0u;
                const uint32_t ck = lane < BK_WIDE_PLAN_WARPS ? S.cand_k[buf][lane] : BK_WIDE_NONE;
                mk = __reduce_max_sync(BK_FULL, key);
                kk = __reduce_min_sync(BK_FULL, key == mk ? ck : BK_WIDE_NONE);
                buf ^= 1u;
            }
            node = node == 0 ? kk + 1 : V.pool[cb + kk];
            depth++;
        }
        // ---- the rest of the walk is warp 0's
        if (warp == 0) {
            uint32_t leaf = node, par = BK_WIDE_NONE, rank = 0;
            bool stop = false;
            if (pending) {                                               // another simulation of a child that is being built
                const bool in = node < BK_WIDE_NC;                       // (its rank sits in move_of until the work kernel runs)
                par = (uint32_t)(in ? S.par[node] : V.parent[node]); rank = in ? S.mv[node] : V.move_of[node];
            } else if (may_expand) {
                const bool in = node < BK_WIDE_NC;
                const uint32_t nl = in ? S.nl[node] : V.n_legal[node], ne = in ? S.ne[node] : V.n_expanded[node];
                uint32_t cb = in ? S.cb[node] : V.child_base[node];
                if (cb == BK_WIDE_NONE) {
                    const uint32_t remaining = (uint32_t)prm.iterations_total - (sims_done + n_items);
                    const uint32_t want = nl < remaining ? nl : remaining;
                    cb = pool_used;
                    if (pool_used + want > L.pool_slots) overflow = 1;
                    pool_used += want;
                }
                if (n_nodes >= L.max_nodes) overflow = 1;
                if (overflow) stop = true;
                else {
                    const uint32_t child = n_nodes++;
                    rank = nl - 1 - ne;
                    __syncwarp();                                        // every lane has read ne / cb before they change
                    if (lane == 0) {
                        V.child_base[node] = cb; V.pool[cb + ne] = child; V.n_expanded[node] = (uint16_t)(ne + 1);
                        V.parent[child] = (int32_t)node; V.visits[child] = 0; V.virt[child] = 0; V.reward[child] = 0;
                        V.child_base[child] = BK_WIDE_NONE; V.n_legal[child] = BK_WIDE_PENDING; V.n_expanded[child] = 0;
                        V.move_of[child] = (uint16_t)rank; V.q[child] = 0.0f; V.inv[child] = 0.0f;
                        if (in) { S.cb[node] = cb; S.ne[node] = (uint16_t)(ne + 1); }
                        if (child < BK_WIDE_NC) {
                            S.par[child] = (int32_t)node; S.vis[child] = 0; S.virt[child] = 0; S.cb[child] = BK_WIDE_NONE;
                            S.nl[child] = BK_WIDE_PENDING; S.ne[child] = 0; S.mv[child] = (uint16_t)rank; S.q[child] = 0.0f; S.inv[child] = 0.0f;
                        }
                    }
                    leaf = child; par = node;
                    __syncwarp();
                }
            }
            if (!stop) {
                if (lane == 0) { V.item_leaf[n_items] = leaf; V.item_parent[n_items] = par; V.item_rank[n_items] = (uint16_t)rank; }
                // one more pending simulation on every node of the path; Q stays the mean of the finished simulations
                for (int32_t v = (int32_t)leaf; v >= 0;) {
                    if (v < BK_WIDE_NC) {
                        const int32_t pend = S.virt[v] + 1;
                        const float iv = __ldg(V.invtab + S.vis[v] + pend);
                        const int32_t up = S.par[v];
                        __syncwarp();
                        if (lane == 0) { S.virt[v] = pend; if (v > 0) S.inv[v] = iv; }
                        v = up;
                    } else {
                        const int32_t pend = V.virt[v] + 1;
                        const float iv = __ldg(V.invtab + V.visits[v] + pend);
                        const int32_t up = V.parent[v];
                        __syncwarp();
                        if (lane == 0) { V.virt[v] = pend; V.inv[v] = iv; }
                        v = up;
                    }
                }
            }
            if (lane == 0) { S.n_nodes = n_nodes; S.pool_used = pool_used; S.overflow = overflow; S.stop = stop ? 1u : 0u; }
        }
        __threadfence_block();
        __syncthreads();
        n_nodes = S.n_nodes; pool_used = S.pool_used; overflow = S.overflow;
        if (S.stop) break;
    }
    // ---- write the round's pending counts and exploration terms of the imaged nodes back
    {
        const uint32_t nc2 = n_nodes < BK_WIDE_NC ? n_nodes : BK_WIDE_NC;
        for (uint32_t v = tid; v < nc2; v += BK_WIDE_PLAN_THREADS) { V.virt[v] = S.virt[v]; if (v > 0) V.inv[v] = S.inv[v]; }
    }
    if (tid == 0) {
        hdr->n_nodes = n_nodes; hdr->pool_used = pool_used; hdr->overflow = overflow;
        hdr->round = hdr->round + 1; hdr->round_items = n_items; hdr->ticket = 0; hdr->sims_base = sims_done;
        hdr->sims_done = sims_done + n_items;
    }
}

// ---- WORK: every warp of the grid takes simulations of tree blockIdx.y from the round's list ---------------------------
template <int POLICY>
__global__ void __launch_bounds__(BK_MCTS_THREADS, POLICY == 0 ? BK_MCTS_CTAS_PER_SM : BK_MCTS_HEUR_CTAS_PER_SM)
k_wide_work(int n_trees, bk_mcts_params prm, unsigned char* __restrict__ workspace) {
    extern __shared__ __align__(16) unsigned char heur_smem[];    // HEUR_SMEM_PER_WARP bytes when POLICY == 1
    const int lane = threadIdx.x & 31;
    const int tree = blockIdx.y;
    if (tree >= n_trees) return;
    const WideLayout L = wide_layout(prm.iterations_total, prm.round_width, prm.max_depth);
    const WideView V = wide_view(workspace + (size_t)tree * L.total, L);
    WideHeader* hdr = V.hdr;
    const uint32_t n_items = hdr->round_items, sims_base = hdr->sims_base;
    const uint32_t root_player = hdr->root_player;
    const int32_t root_score = hdr->root_score;
    const uint64_t tree_id = prm.tree_offset + (uint64_t)tree;
    for (;;) {
        uint32_t t0 = 0;
        if (lane == 0) t0 = atomicAdd(&hdr->ticket, 1u);
        const uint32_t i = uni(t0);
        if (i >= n_items) break;
        const uint32_t leaf = V.item_leaf[i], par = V.item_parent[i];
        WarpState s;
        if (par != BK_WIDE_NONE) {                                       // expansion (mcts_agent.py:113-145)
            load_state(V.states + par, lane, s);
            const uint32_t p = s.to_move;
            uint32_t A, C; make_masks(s, p, lane, A, C);
            const Totals tot = gen_totals<1>(A, C, sel4(s.used, p), lane);
            int o, row, col;
            const uint32_t id = select_move(tot, V.item_rank[i], A, C, lane, o, row, col);
            place_move(s, p, o, row, col, lane);
            uint32_t A2, C2; make_masks(s, s.to_move, lane, A2, C2);
            const Totals tc = gen_totals<1>(A2, C2, sel4(s.used, s.to_move), lane);
            store_state(V.states + leaf, lane, s);
            if (lane == 0) { V.n_legal[leaf] = (uint16_t)tc.n; V.move_of[leaf] = (uint16_t)id; }
        } else {
            load_state(V.states + leaf, lane, s);
        }
        // ---- simulation (mcts_agent.py:470-554): simulation number j draws from its own Philox blocks
        long long r = 0;
        if (prm.reward_mode == 2) {
            r = 1;                                                       // the dummy evaluator's 0.5, counted as one half
        } else {
            const uint32_t j = sims_base + i;
            uint32_t draws = 0, blk_of = 0;
            PhiloxBlock blk = philox_block(prm.seed, tree_id, j << 5, BK_WIDE_STREAM);
            auto next_draw = [&]() -> uint32_t {
                if ((draws >> 2) != blk_of) { blk_of = draws >> 2; blk = philox_block(prm.seed, tree_id, (j << 5) | blk_of, BK_WIDE_STREAM); }
                return philox_word(blk, draws++);
            };
            const uint32_t me = prm.reward_mode ? root_player : s.to_move;
            const int32_t before = prm.reward_mode ? root_score : board_score_of(s, me);
            const int limit = prm.max_rollout_moves > 0 ? prm.max_rollout_moves : 96;    // a game has at most 84 placements
            for (int ply = 0; ply < limit; ply++) {
                const uint32_t p = s.to_move;
                if (POLICY == BK_POLICY_HEURISTIC) {
                    uint32_t id = 0, idx = 0;
                    const uint32_t before_draws = draws;
                    const uint32_t rnd = next_draw();
                    if (heuristic_pick(s, p, rnd, 0u, lane, false, id, idx) == 0) { draws = before_draws; break; }
                    place_move(s, p, (int)(id / 400u), (int)((id % 400u) / 20u), (int)(id % 20u), lane);
                    continue;
                }
                uint32_t A, C; make_masks(s, p, lane, A, C);
                const Totals tot = gen_totals<1>(A, C, sel4(s.used, p), lane);
                if (tot.n == 0) break;
                const uint32_t rnd = next_draw();
                int o, row, col;
                select_move(tot, __umulhi(rnd, tot.n), A, C, lane, o, row, col);
                place_move(s, p, o, row, col, lane);
            }
            r = (long long)(board_score_of(s, me) - before);
        }
        // ---- back-propagation (mcts_agent.py:572-582) with integer atomics: any order gives the same sums
        if (lane == 0) {
            for (int32_t v = (int32_t)leaf; v >= 0; v = V.parent[v]) {
                atomicAdd(&V.visits[v], 1);
                atomicAdd(reinterpret_cast<unsigned long long*>(&V.reward[v]), (unsigned long long)r);
            }
        }
        __syncwarp();
    }
}

// ---- root statistics by ascending legal-move rank (child k of the root is rank n_legal-1-k) ----------------------------
__global__ void __launch_bounds__(32)
k_wide_finish(int n_trees, bk_mcts_params prm, unsigned char* __restrict__ workspace, int max_root_moves,
              int32_t* __restrict__ out_visits, long long* __restrict__ out_reward, uint32_t* __restrict__ out_counts,
              uint16_t* __restrict__ out_moves, uint32_t* __restrict__ out_status, uint32_t* __restrict__ out_done) {
    const int lane = threadIdx.x & 31;
    const int tree = blockIdx.x;
    if (tree >= n_trees) return;
    const WideLayout L = wide_layout(prm.iterations_total, prm.round_width, prm.max_depth);
    const WideView V = wide_view(workspace + (size_t)tree * L.total, L);
    const uint32_t nl0 = V.n_legal[0], ne0 = V.n_expanded[0], cb0 = V.child_base[0];
    if (lane == 0) {
        if (out_status) out_status[tree] = V.hdr->overflow;
        if (out_counts) out_counts[tree] = nl0;
        if (out_done) out_done[tree] = V.hdr->sims_done;
    }
    for (uint32_t rank = lane; rank < (uint32_t)max_root_moves; rank += 32) {
        int32_t v = 0; long long w = 0; uint16_t m = 0xFFFF;
        if (rank < nl0) {
            const uint32_t k = nl0 - 1 - rank;
            if (k < ne0) { const uint32_t ch = V.pool[cb0 + k]; v = V.visits[ch]; w = V.reward[ch]; m = V.move_of[ch]; }
        }
        const size_t oi = (size_t)tree * max_root_moves + rank;
        if (out_visits) out_visits[oi] = v;
        if (out_reward) out_reward[oi] = w;
        if (out_moves) out_moves[oi] = m;
    }
}
