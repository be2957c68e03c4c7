/*
 * blokus_b200.h -- C ABI of the B200-native Blokus simulation core (libblokus_b200.so).
 *
 * The reference (TGALLOWAY1/ReinforcementLearning_Blokus) is pure Python and has no FFI; its
 * seams are Python call signatures.  Every entry point below names the reference call it
 * replaces (paths relative to the reference root).  Conventions:
 *   - plain pointers and sizes only; pointers are DEVICE pointers unless the name ends in
 *     `_host` or the comment says "host";
 *   - every function returns 0 on success or a negative bk_status; nothing throws;
 *     bk_last_error() returns a thread-local message for the last failure;
 *   - the caller owns every buffer (the library itself keeps only constant tables and a small pool of
 *     per-launch work counters); `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream); calls on distinct streams may run concurrently;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     BK_ERR_CUDA.
 *
 * Packed game state (bk_state, 256 bytes, 256-byte aligned in device arrays):
 *   bits[p][w]   400-bit board of player p (0 RED,1 BLUE,2 YELLOW,3 GREEN); bit row*20+col,
 *                word w holds bits 64w..64w+63 -- byte-identical to
 *                int(board.player_bits[Player(p+1)]).to_bytes(56,'little') (engine/bitboard.py:13-22)
 *   used[p]      bit (piece_id-1) set when the piece is spent (engine/board.py:63)
 *   to_move      board.current_player.value-1                  (engine/board.py:66)
 *   first_mask   bit p set while board.player_first_move[p]    (engine/board.py:64)
 *   move_count   board.move_count                              (engine/board.py:67)
 *
 * Move ids: uint16 id = global_orientation(0..90)*400 + anchor_row*20 + anchor_col, the anchor being
 * the bounding-box top-left as in engine/move_generator.py:90-103.  Ascending id order equals the
 * (piece, orientation, row, col) order of the reference's naive generator
 * (engine/move_generator.py:153-259).  BK_PASS (0xFFFF) marks a pass / padding.
 */
#ifndef BLOKUS_B200_H
#define BLOKUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_BOARD 20
#define BK_CELLS 400
#define BK_WORDS 7
#define BK_PIECES 21
#define BK_ORIENTS 91
#define BK_PASS 0xFFFFu
#define BK_MAX_TURNS 96        /* trace slots per playout: <= 21 rounds * 4 + 4 closing passes = 88 */
#define BK_BITMAP_WORDS (BK_ORIENTS * BK_BOARD) /* legal-anchor bitmap: one 20-bit row word per (orientation,row) */
#define BK_ZOBRIST_ENTRIES 2088 /* 20*20*5 + 4 + 4*21, reference order (mcts/zobrist.py:38-68) */

typedef struct bk_state {
    uint64_t bits[4][BK_WORDS];
    uint32_t used[4];
    uint8_t  to_move;
    uint8_t  first_mask;
    uint16_t move_count;
    uint8_t  reserved[12];
} bk_state;

enum bk_status {
    BK_OK = 0,
    BK_ERR_CUDA = -1,        /* CUDA runtime error or no device */
    BK_ERR_ARG = -2,         /* bad argument */
    BK_ERR_NOT_INIT = -3,    /* bk_init not called on this device */
    BK_ERR_CAPACITY = -4     /* a caller-supplied buffer is too small */
};

enum bk_policy {
    BK_POLICY_UNIFORM = 0,   /* agents/random_agent.py:33-50 (uniform index into the legal list) */
    BK_POLICY_HEURISTIC = 1  /* agents/heuristic_agent.py:39-66 (softmax over move scores) */
};

/* ---- lifecycle -------------------------------------------------------------------------- */

/* Create the library's context for `device` (piece tables in its constant memory, per-launch work counters); the
 * caller's current device is left unchanged and other devices' contexts are untouched.  Every other entry point works
 * on the CALLER'S CURRENT DEVICE (the one its stream and buffers belong to) and returns BK_ERR_NOT_INIT if bk_init was
 * not called for it.  Replaces the import-time table construction of engine/pieces.py:496-513 and
 * LegalMoveGenerator.__init__ (engine/move_generator.py:112-128). */
int bk_init(int device);
const char* bk_last_error(void);
/* Library build id string (arch, build flags). */
const char* bk_version(void);

/* ---- K1: legal-move generation  (engine/move_generator.py:130-151, 261-559) --------------- */

/* Count legal moves of `players[i]` (NULL: the side to move) in states[i].
 * counts[i] = len(get_legal_moves(board, player)). */
int bk_movegen_count(const bk_state* states, const uint8_t* players, int n,
                     uint32_t* counts, void* stream);

/* Exclusive prefix sum of counts[0..n) into offsets[0..n]; offsets[n] = total. */
int bk_exclusive_scan_u32(const uint32_t* counts, int n, uint32_t* offsets, void* stream);

/* Write the ascending move ids of state i to move_ids[offsets[i] .. offsets[i+1]). */
int bk_movegen_emit(const bk_state* states, const uint8_t* players, int n,
                    const uint32_t* offsets, uint16_t* move_ids, void* stream);

/* One-kernel variant with a fixed capacity per state: ids of state i go to move_ids[i*cap ..], at most
 * cap of them; counts[i] is the full count (so counts[i] > cap tells the caller the list was cut). */
int bk_movegen_fixed(const bk_state* states, const uint8_t* players, int n, uint32_t cap,
                     uint32_t* counts, uint16_t* move_ids, void* stream);

/* ONE board with HOST buffers -- LegalMoveGenerator.get_legal_moves(board, player) for a single position
 * (engine/move_generator.py:130-151; the protocol of benchmarks/benchmark_move_generation.py:136-224): copies the 256-byte
 * state up, generates the side to move's ascending id list, copies it back and waits for the stream.  state_host: the
 * packed state with `to_move` set to the player asked for; state_dev (256 B) and list_dev (capacity*2 + 8 bytes) are the
 * caller's device staging; list_host receives capacity uint16 ids followed by the uint32 count (a count above capacity means
 * the list was truncated).  Page-locked host buffers make the copies asynchronous. */
int bk_legal_moves_host(const bk_state* state_host, bk_state* state_dev, void* list_dev, void* list_host, int capacity,
                        void* stream);

/* One-pass CSR: after bk_movegen_fixed (ids in rows of `cap` slots, full lengths in counts) and bk_exclusive_scan_u32 of
 * min(counts, cap), copy every row's ids to its CSR range move_ids[offsets[i] .. offsets[i+1]).  Rows longer than cap
 * must be regenerated with bk_movegen_emit (the host wrapper does that when it happens). */
int bk_compact_rows(const uint16_t* slots, uint32_t cap, const uint32_t* offsets, int n, uint16_t* move_ids, void* stream);

/* Legal-anchor bitmap: bitmap[i][o*20+row] has bit col set iff move (o,row,col) is legal.  `bitmap` must be 16-byte
 * aligned (each state's 7 280 bytes leave as 16-byte stores). */
int bk_movegen_bitmap(const bk_state* states, const uint8_t* players, int n,
                      uint32_t* bitmap, void* stream);

/* Per-piece placement counts P_i of players[i] (NULL: side to move): piece_counts[i][pid-1] = number of legal moves
 * that use piece pid -- the movegen-bound input of compute_player_mobility_metrics (engine/mobility_metrics.py:38-131,
 * used by the snapshot features analytics/winprob/features.py:131-193).  anchor_max[i] (optional) = the largest
 * number of legal moves that share one anchor cell (numerator of anchorTop1Share). */
int bk_movegen_piece_counts(const bk_state* states, const uint8_t* players, int n, uint16_t* piece_counts /*[n][21]*/,
                            uint16_t* anchor_max /*[n] or NULL*/, void* stream);

/* Win-probability snapshot feature rows: features[i][p][0..33] = the 34 columns of SNAPSHOT_FEATURE_COLUMNS
 * (analytics/winprob/features.py:23-58) for player p of states[i], as build_snapshot_runtime_context (:103-128) +
 * extract_player_snapshot_features (:131-193) compute them -- the input of the arena's snapshot dataset
 * (analytics/tournament/arena_runner.py:522-575) and of LearnedWinProbabilityEvaluator (mcts/learned_evaluator.py:79-95).
 * frontiers: uint64[n][4][7] bitboards of the reference's incrementally maintained frontier sets (Board.get_frontier,
 * engine/board.py:247-259), or NULL to use the sets a freshly restored board has (_compute_full_frontier, :261-313,
 * plus the start corner before the first move).  turn_index: int32[n] or NULL (= move_count, as the evaluator passes). */
#define BK_SNAPSHOT_FEATURES 34
int bk_snapshot_features(const bk_state* states, const uint64_t* frontiers, const int32_t* turn_index, int n,
                         int max_turns, double* features /*[n][4][34]*/, void* stream);

/* Replay n recorded games (traces[g][t] = move id of turn t, BK_PASS = pass; start = NULL for the empty board) and
 * capture the feature rows where analytics/tournament/arena_runner.py:625-650,698 captures them: before the first turn
 * and after each placement whose move_count equals a checkpoint ply, with turn_index = turns played so far (passes
 * included) and the frontier sets maintained by the rule of engine/board.py:315-367.  checkpoints_host: HOST array of
 * 1..32 distinct plies.  features[g][k][p][34], turn_index[g][k] (-1: game g never reached checkpoint k); optional
 * snap_states[g][k] and snap_frontiers[g][k][4][7] receive the captured boards. */
int bk_replay_snapshots(const bk_state* start, const uint16_t* traces, int n, int n_turns,
                        const int32_t* checkpoints_host, int n_checkpoints, int max_turns, double* features,
                        int32_t* turn_index, bk_state* snap_states, uint64_t* snap_frontiers, void* stream);

/* has_legal_moves for all four players (engine/move_generator.py:961-1054): bit p of out[i]. */
int bk_has_moves(const bk_state* states, int n, uint8_t* out_mask, void* stream);

/* is_move_legal (engine/move_generator.py:912-946): out[i] = 1 if move_ids[i] is legal for players[i]. */
int bk_is_move_legal(const bk_state* states, const uint8_t* players, const uint16_t* move_ids,
                     int n, uint8_t* out, void* stream);

/* Board.can_place_piece (engine/board.py:136-220) for an arbitrary cell set: cell_masks[i] is a 400-bit
 * mask (7 words, bit row*20+col) of the cells to cover; out[i] = 1 if players[i] may place exactly those
 * cells (all empty, none edge-adjacent to own cells, one on an own corner -- or on the start corner for
 * the first move).  The piece inventory is not consulted, as in the reference. */
int bk_can_place_mask(const bk_state* states, const uint8_t* players, const uint64_t* cell_masks, int n,
                      uint8_t* out, void* stream);

/* ---- step: BlokusGame.make_move + caller-side pass (engine/game.py:57-180, board.py:515-560) */

/* For each i: move_ids[i]==BK_PASS advances to_move only (engine/board.py:557-560); otherwise the
 * move is checked for players[i] (NULL: the side to move) and, if legal, applied (cells, used, first-move flag,
 * move_count, to_move).  ok[i] = 1 applied / passed, 0 rejected (state unchanged), mirroring
 * make_move's bool.  alive[i] (optional) = has-moves mask after the step (game over iff 0,
 * engine/game.py:182-214). */
int bk_step(bk_state* states, const uint8_t* players, const uint16_t* move_ids, int n,
            uint8_t* ok, uint8_t* alive, void* stream);

/* ---- K3: scores (engine/board.py:562-577, engine/game.py:284-349) ------------------------- */
int bk_scores(const bk_state* states, int n, int32_t* game_scores /*[n][4]*/,
              int32_t* board_scores /*[n][4]*/, void* stream);

/* ---- K2: playouts to terminal score (loop of analytics/tournament/arena_runner.py:653-695) - */

/* Game i (global index game_offset+i) starts from start_states[i] (NULL: the empty board) and plays
 * turns first_turn, first_turn+1, ... (passes count as turns): turn t of game g draws
 * r = word t%4 of philox4x32_10(key=seed, ctr=(g_lo,g_hi,t/4,0)) and plays the ((uint64)r*n_legal)>>32-th move of the
 * ascending list (policy BK_POLICY_UNIFORM; with BK_POLICY_HEURISTIC the same word r is the categorical draw over the
 * softmax of HeuristicAgent's move scores, cumulative order as documented in csrc/heuristic.cuh and restated by the oracle);
 * a player with no move passes; the game stops once no player can move or when
 * t reaches max_turns (0 = BK_MAX_TURNS).  A game played in segments [0,a), [a,b), ... through
 * final_states is the same game as one call.  Outputs (each may be NULL): game_scores/board_scores [n][4],
 * trace [n][BK_MAX_TURNS] move ids by turn (BK_PASS = pass/pad; slots < first_turn are left untouched),
 * choice [n][BK_MAX_TURNS] rank of each chosen move in the ascending legal-move list, final_states [n]. */
int bk_playout(const bk_state* start_states, int n, uint64_t seed, uint64_t game_offset,
               int policy, int first_turn, int max_turns, int32_t* game_scores, int32_t* board_scores,
               uint16_t* trace, uint16_t* choice, bk_state* final_states, void* stream);

/* Arena match-ups (analytics/tournament/arena_runner.py:578-760 with per-seat agents): as bk_playout, but the
 * policy of player p in game i is byte p of seat_policies[i] (bk_policy values), so RandomAgent and
 * HeuristicAgent seats can be mixed and rotated per game in one batch. */
int bk_playout_seats(const bk_state* start_states, int n, uint64_t seed, uint64_t game_offset,
                     const uint32_t* seat_policies, int first_turn, int max_turns, int32_t* game_scores,
                     int32_t* board_scores, uint16_t* trace, uint16_t* choice, bk_state* final_states,
                     void* stream);

/* ---- K4: Zobrist keys and transposition table (mcts/zobrist.py:20-99, 155-220) ------------ */

/* keys[i] = ZobristHash.hash_board(board_i)  (mcts/zobrist.py:70-99).  `table` is a DEVICE array of
 * BK_ZOBRIST_ENTRIES keys in the reference's draw order: position_player_hashes[20][20][5],
 * player_turn_hashes[4], piece_used_hashes[4][21] (mcts/zobrist.py:38-68). */
int bk_zobrist_hash(const bk_state* states, int n, const uint64_t* table, uint64_t* keys, void* stream);

typedef struct bk_tt bk_tt;  /* open-addressing device table {key -> reward} */
int bk_tt_create(bk_tt** out, uint32_t log2_slots);
int bk_tt_clear(bk_tt* tt, void* stream);
int bk_tt_destroy(bk_tt* tt);
/* TranspositionTable.get / put (mcts/zobrist.py:171-202); found[i]=1 and rewards[i] set on a hit. */
int bk_tt_get(bk_tt* tt, const uint64_t* keys, int n, double* rewards, uint8_t* found, void* stream);
int bk_tt_put(bk_tt* tt, const uint64_t* keys, const double* rewards, int n, void* stream);
/* host out[4]: [0]=slots, [1]=entries, [2]=hits and [3]=probes of the searches that used the table
 * (TranspositionTable.hit_count / access_count, mcts/zobrist.py:181-185; reset by bk_tt_clear) */
int bk_tt_stats(bk_tt* tt, uint64_t* out_host, void* stream);
/* TranspositionTable.put's eviction (mcts/zobrist.py:196-200): remove the n_evict entries inserted first. */
int bk_tt_evict_oldest(bk_tt* tt, uint64_t n_evict, void* stream);

/* ---- N1: HeuristicAgent move scores (agents/heuristic_agent.py:68-198) ---------------------- */

/* scores[k] for every move_ids[k], k in [offsets[i], offsets[i+1]), evaluated on states[i] for players[i]
 * (NULL: side to move): 1.0*size + 2.0*corner_creation - 1.5*edge + 0.5*centre in float64, bit-identical to
 * HeuristicAgent._evaluate_move.  The softmax draw over them stays with the caller (:222-244, :65). */
int bk_heuristic_scores(const bk_state* states, const uint8_t* players, const uint32_t* offsets,
                        const uint16_t* move_ids, int n, double* scores, void* stream);

/* The same with HeuristicAgent.set_weights (agents/heuristic_agent.py:268-282): weights_host = HOST array
 * {piece_size, corner_creation, edge_avoidance, center_preference}; products and sums in the reference's order. */
int bk_heuristic_scores_weighted(const bk_state* states, const uint8_t* players, const uint32_t* offsets,
                                 const uint16_t* move_ids, int n, const double* weights_host, double* scores, void* stream);

/* ---- K5: UCT search (mcts/mcts_agent.py:304-582; flat variant agents/fast_mcts_agent.py:112-338) -- */

typedef struct bk_mcts_params {
    int32_t  iterations;        /* iterations to run in THIS call (MCTSAgent.iterations, mcts_agent.py:204) */
    int32_t  iterations_total;  /* capacity of the whole search (>= sum over resumed calls); sizes the workspace */
    int32_t  max_rollout_moves; /* mcts_agent.py:218 (50); 0 = roll out until the side to move has no move */
    int32_t  rollout_policy;    /* bk_policy */
    int32_t  max_depth;         /* 0 = full UCT tree; 1 = flat bandit over the root moves (FastMCTSAgent) */
    int32_t  ucb_variant;       /* 0: c*sqrt(ln N/n) (mcts_agent.py:86); 1: c*sqrt(2 ln N/n) (fast_mcts_agent.py:52) */
    int32_t  resume;            /* 0: build a fresh tree from the root; 1: continue the trees in `workspace` */
    int32_t  reward_mode;       /* 0: gain of the player to move at the leaf, from the leaf (mcts_agent.py:470-521);
                                   1: gain of the ROOT player measured from the root position (flat bandit);
                                   2: no rollout, every leaf is worth the constant 0.5 of the reference's dummy learned
                                      evaluator (mcts_agent.py:427-451, learned_evaluator.py:51-55); rewards are then
                                      counted in halves (root_reward = number of halves);
                                   3: no rollout, the leaf is worth the learned win probability of its player to move
                                      (mcts_agent.py:427-451; LearnedWinProbabilityEvaluator with a pairwise logistic
                                      model, mcts/learned_evaluator.py:97-143): the 4 x 34 snapshot features of the leaf
                                      are computed in the kernel and fed to leaf_model.  Rewards are doubles:
                                      root_reward then holds IEEE-754 bit patterns of the reward sums;
                                   4: rollout as in mode 0 plus potential-based shaping with the same learned model
                                      (leaf_model, shaping_weight, shaping_gamma, potential_mode); rewards are doubles */
    int32_t  leaf_batch;        /* 1 (or 0): one rollout per iteration, one warp per tree -- the reference's search.
                                   2..16: leaf-batched rollouts -- a CTA of that many warps owns the tree; after each
                                   selection/expansion every warp rolls out from the new leaf with its own random stream
                                   and the rewards are backed up together (visits += leaf_batch); a transposition
                                   table then caches the SUM of a leaf's rewards. */
    int32_t  tt_tree_base;      /* transposition-table namespace of tree k = tt_tree_base + k (independent of tree_offset):
                                   an agent keeps it constant over its searches so that a later search finds what an
                                   earlier one cached (mcts_agent.py:337-339, 418-434); concurrently searched trees get
                                   different namespaces and never read each other's entries.  Namespace 0 = plain Zobrist
                                   keys, the reference's single table. */
    double   exploration;       /* 1.414 */
    uint64_t seed;
    uint64_t tree_offset;       /* global index of tree 0 (root-parallel sharding across GPUs) */
    const double*   leaf_model;     /* reward_mode 3: DEVICE pointer to 35 doubles, weights of the 34 snapshot feature
                                       differences + bias of a "pairwise_logreg" win-probability model */
    const uint64_t* root_frontiers; /* reward_mode 3: DEVICE [n_roots][4][7] frontier sets of the roots (Board.get_frontier,
                                       engine/board.py:247-259), or NULL for the sets of a freshly restored board */
    int32_t  leaf_max_turns;        /* reward_mode 3: max_turns of the evaluator's progress features (2500) */
    int32_t  leaf_model_kind;       /* 0: leaf_model = 35 doubles (logistic); 1: flattened gradient-boosted phase models
                                       ("pairwise_gbt_phase"; layout in csrc/snapshot.cuh / winprob.flatten_gbt_phase_models) */
    double   progressive_bias_weight; /* reward_mode 3: w of the selection term w * prior/(1 + n) (mcts_agent.py:88, 453-468),
                                         prior(child) = V(child board, parent player) - V(parent board, parent player);
                                         0 = off */
    double   shaping_weight;          /* reward_mode 4: w and gamma of the potential-based shaping term added to TRUNCATED */
    double   shaping_gamma;           /*   rollouts, w * (gamma * Phi(final) - Phi(initial)) (mcts_agent.py:486-551) */
    int32_t  potential_mode;          /* reward_mode 4: Phi = clamped win probability (0, "prob") or its logit (1, "logit") */
    int32_t  round_width;             /* bk_mcts_wide_search only: simulations planned per round (1..4096) */
} bk_mcts_params;

/* Bytes of device workspace per tree for a search of params->iterations_total iterations (host call). */
int64_t bk_mcts_workspace_bytes(const bk_mcts_params* params);

/* Run n_trees independent UCT searches, one warp each (a CTA of params->leaf_batch warps each when leaf_batch > 1);
 * tree k searches roots[k % n_roots] (side to move = the root player) with random stream (seed, tree_offset+k).  `tt` (may be NULL) caches rollout rewards by
 * Zobrist key of `zobrist_table` (device, BK_ZOBRIST_ENTRIES; required when tt != NULL).  Outputs per tree,
 * indexed by the rank of the move in the root's ascending legal-move list:
 *   root_visits [n_trees][max_root_moves] int32, root_reward [n_trees][max_root_moves] int64 (rollout
 *   rewards are integers), root_move_ids [n_trees][max_root_moves] (BK_PASS where not expanded),
 *   root_counts [n_trees] = number of legal root moves, status [n_trees] = 1 if the tree ran out of workspace. */
int bk_mcts_search(const bk_state* roots, int n_roots, int n_trees, const bk_mcts_params* params,
                   bk_tt* tt, const uint64_t* zobrist_table, void* workspace, int64_t workspace_bytes,
                   int max_root_moves, int32_t* root_visits, int64_t* root_reward, uint32_t* root_counts,
                   uint16_t* root_move_ids, uint32_t* status, void* stream);

/* ---- K6: one tree over the whole GPU (csrc/mcts_wide.cuh) ------------------------------------------------------------
 * The searches above run the reference's sequential iteration, one warp (or CTA) per tree.  A caller that needs ONE move
 * fast -- FastMCTSAgent.think (agents/fast_mcts_agent.py:112), POST /think (engine-service/app.py:53-72), a single
 * MCTSAgent.select_action (mcts/mcts_agent.py:304) -- uses this round-synchronous search instead: per round one warp
 * plans params->round_width simulations by UCB1 walks with virtual loss (expanding in the reference's pop-last order), the
 * whole grid executes them (expansion + Philox rollout of simulation number j + integer back-propagation), repeat.
 * Deterministic: the result depends on (seed, tree_offset + k, round_width and the sequence of calls) only.
 * params: iterations = simulations to add in THIS call, iterations_total = capacity, resume, max_depth, ucb_variant,
 * reward_mode 0/1/2, rollout_policy, max_rollout_moves (<= 96), exploration, seed, tree_offset, round_width.  No
 * transposition table and no learned modes here.  The call returns after the search has finished (it synchronises the
 * stream to learn how many rounds were needed).  Outputs as bk_mcts_search, plus sims_done [n_trees]. */
int64_t bk_mcts_wide_workspace_bytes(const bk_mcts_params* params);
int bk_mcts_wide_search(const bk_state* roots, int n_roots, int n_trees, const bk_mcts_params* params,
                        void* workspace, int64_t workspace_bytes, int max_root_moves, int32_t* root_visits,
                        int64_t* root_reward, uint32_t* root_counts, uint16_t* root_move_ids, uint32_t* status,
                        uint32_t* sims_done, void* stream);

/* ---- root-parallel merge over NVLink peer memory (the one exchange step of a root-parallel search) ------
 * The reference searches one tree per move (mcts/mcts_agent.py:304-358); the multi-GPU path builds root-parallel trees and
 * sums their root statistics.  Instead of ncclAllReduce(sum) after the search, every rank runs ONE kernel that sums its
 * own trees, adds the sums into EVERY rank's window with 64-bit atomics on peer memory (NVLink / NVSwitch), signals, waits
 * until all ranks have arrived and reads its own window: the merged statistics are on every rank when the kernel ends.
 * A window is created by its owner (cudaMalloc + cudaIpcGetMemHandle), its 64-byte handle travels to the other processes
 * by any host channel, and they map it with bk_p2p_window_open.
 *   bk_p2p_window_create : `capacity` = largest number of root moves (<= 2048); *window = device pointer, handle_host[64]
 *   bk_root_merge_p2p    : root_visits [n_trees][stride] int32 (or int64 when visits_are_int64: sums the caller has already
 *                          taken over many trees), root_reward [n_trees][stride] int64 (outputs of bk_mcts_search /
 *                          bk_mcts_wide_search; one CTA adds up the trees, so keep n_trees small), n = number of root moves; windows_host = host array of
 *                          `world` device pointers in rank order (own window at [rank]); epoch = 0, 1, 2, ... the same on
 *                          every rank, +1 per merge; merged [2][n] int64 (visits, reward); status = 0, or 1 if the peers
 *                          did not arrive within timeout_ms (<= 0: 20 s; merged is then undefined).  One process per GPU. */
int bk_p2p_window_create(int capacity, void** window, unsigned char* handle_host);
int bk_p2p_window_open(const unsigned char* handle_host, void** peer_window);
int bk_p2p_window_close(void* peer_window);
int bk_p2p_window_destroy(void* window);
int bk_root_merge_p2p(const void* root_visits, int visits_are_int64, const int64_t* root_reward, int n_trees, int stride, int n, int capacity,
                      int rank, int world, void* const* windows_host, uint64_t epoch, int timeout_ms, int64_t* merged,
                      int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BLOKUS_B200_H */
