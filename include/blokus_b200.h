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
 *   - the caller owns every buffer; `stream` is a cudaStream_t passed as void* (NULL = the
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

/* Select `device`, upload the piece tables.  Replaces the import-time table construction of
 * engine/pieces.py:496-513 and LegalMoveGenerator.__init__ (engine/move_generator.py:112-128). */
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

/* Legal-anchor bitmap: bitmap[i][o*20+row] has bit col set iff move (o,row,col) is legal. */
int bk_movegen_bitmap(const bk_state* states, const uint8_t* players, int n,
                      uint32_t* bitmap, void* stream);

/* has_legal_moves for all four players (engine/move_generator.py:961-1054): bit p of out[i]. */
int bk_has_moves(const bk_state* states, int n, uint8_t* out_mask, void* stream);

/* is_move_legal (engine/move_generator.py:912-946): out[i] = 1 if move_ids[i] is legal for players[i]. */
int bk_is_move_legal(const bk_state* states, const uint8_t* players, const uint16_t* move_ids,
                     int n, uint8_t* out, void* stream);

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

/* Game i (global index game_offset+i) starts from start_states[i] (NULL: the empty board) and is
 * played to the end: turn t (passes count) of game g draws r = philox4x32_10(key=seed,
 * ctr=(g_lo,g_hi,t,0)).x and plays the ((uint64)r*n_legal)>>32-th move of the ascending list; a
 * player with no move passes; the game stops once no player can move or after max_turns turns
 * (0 = BK_MAX_TURNS).  Outputs (each may be NULL): game_scores/board_scores [n][4],
 * trace [n][BK_MAX_TURNS] move ids (BK_PASS = pass/pad), choice [n][BK_MAX_TURNS] chosen indices,
 * final_states [n]. */
int bk_playout(const bk_state* start_states, int n, uint64_t seed, uint64_t game_offset,
               int policy, int max_turns, int32_t* game_scores, int32_t* board_scores,
               uint16_t* trace, uint16_t* choice, bk_state* final_states, void* stream);

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
/* host out: [0]=slots, [1]=occupied */
int bk_tt_stats(bk_tt* tt, uint64_t* out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BLOKUS_B200_H */
