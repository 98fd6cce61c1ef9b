/*
 * oracle/xq_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's Xiangqi rules engine and MCTS loop
 * (reference files: chess_env.py, self_play.py, neural_network.py; cited per
 * function in xq_oracle.c).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (chinesechessai_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED — every function here is differentially checked
 * against the imported Python reference by tests/golden/gen_golden.py (run in
 * the authoring container, where /root/reference exists); the resulting
 * traces are committed under tests/golden/ and re-checked on every
 * `pytest -m "not gpu"` run by tests/test_oracle_golden.py.
 */
#ifndef XQ_ORACLE_H
#define XQ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XQO_ROWS 10
#define XQO_COLS 9
#define XQO_NSQ 90
#define XQO_MAX_MOVES 128
#define XQO_HIST_CAP 1024
#define XQO_WINNER_NONE 2 /* Python None */

/* end_reason classes (chess_env.py:297,359,366,373,381,389,397,404) */
enum {
  XQO_REASON_NONE = 0,
  XQO_REASON_KING_CAPTURE = 1,    /* "{mover}吃掉对方将帅"        */
  XQO_REASON_CHECKMATE = 2,       /* "将死{side to move}"         */
  XQO_REASON_REPETITION = 3,      /* "三次重复局面判和"           */
  XQO_REASON_FIFTY = 4,           /* "50回合无吃子判和"           */
  XQO_REASON_STALEMATE = 5,       /* "困毙{side to move}"         */
  XQO_REASON_PERPETUAL_CHECK = 6, /* "长将判负({side to move})"   */
  XQO_REASON_PERPETUAL_CHASE = 7, /* never produced (chess_env.py:674) */
  XQO_REASON_MOVE_CAP = 8         /* "超过{move_count}步判和"     */
};

typedef struct {
  int8_t board[XQO_NSQ]; /* row-major r*9+c, codes of config.py:66-74 */
  int8_t pad_[6];
  int32_t player;     /* +1 red / -1 black (chess_env.py:62) */
  int32_t move_count; /* chess_env.py:63 */
  int32_t winner;     /* 1 / -1 / 0 / XQO_WINNER_NONE (chess_env.py:64) */
  int32_t reason;     /* XQO_REASON_* */
  int32_t red_king;   /* cached square or -1 for None (chess_env.py:27,44) */
  int32_t black_king; /* chess_env.py:28,57 */
  int32_t no_capture; /* chess_env.py:21 */
  int32_t consecutive_checks; /* chess_env.py:24 */
  int32_t pos_len;            /* len(position_history) */
  int32_t check_len;          /* len(check_history)    */
  int32_t overflow;           /* set if a history exceeded XQO_HIST_CAP */
  int32_t pad2_;
  uint64_t pos_hist[XQO_HIST_CAP];
  uint8_t check_hist[XQO_HIST_CAP];
} xqo_state;

typedef struct {
  double reward;
  int32_t reward_is_int; /* 1 if the reference returns a Python int */
  int32_t done;
} xqo_step_result;

/* --- rules (chess_env.py) ------------------------------------------------ */
void xqo_reset(xqo_state *s);
uint64_t xqo_position_hash(const int8_t *board, int player);
int xqo_legal_moves(xqo_state *s, int16_t *moves /* [XQO_MAX_MOVES] */);
int xqo_pseudo_count(xqo_state *s); /* candidates before the suicide filter */
int xqo_is_in_check(const xqo_state *s, int player);
int xqo_kings_facing(const xqo_state *s);
void xqo_make_move(xqo_state *s, int move /* from*90+to */,
                   xqo_step_result *out);

/* --- counter-based move pick shared by CPU and GPU playouts --------------- */
void xqo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]);
int xqo_pick_move(const xqo_state *s, const int16_t *moves, int n,
                  uint64_t seed, uint32_t game_id, uint32_t ply,
                  int capture_bias /* 0..256 */);

typedef struct {
  int32_t plies;
  int32_t winner;
  int32_t reason;
  int32_t max_legal;
  double reward_sum;   /* sequential float64 sum of step rewards */
  uint64_t digest;     /* running digest of (move list, board, flags) per ply */
  uint64_t final_hash; /* xqo_position_hash of the final board/side */
} xqo_playout_result;

/* trace buffers may be NULL; sized [max_plies] (moves: [max_plies][128]) */
void xqo_playout(xqo_state *s, uint64_t seed, uint32_t game_id, int max_plies,
                 int capture_bias, xqo_playout_result *res,
                 int16_t *trace_moves, int16_t *trace_n, int16_t *trace_pick,
                 double *trace_reward, uint8_t *trace_flags,
                 int8_t *trace_boards /* [max_plies][90], board after ply */);

/* n_games playouts from the initial position on n_threads pthreads; returns
 * total plies.  results: [n_games]. */
int64_t xqo_playout_many(int n_games, uint32_t first_game_id, uint64_t seed,
                         int max_plies, int capture_bias, int n_threads,
                         xqo_playout_result *results);

/* Move choice of self_play_game (self_play.py:219-243) with the counter-based uniform the CUDA
 * engine uses (philox ctr=(game_id, ply, 1, 0)); returns the index into the visit list. */
int xqo_sample_move(const int32_t *visits, int n, double temperature, uint64_t seed,
                    uint32_t game_id, uint32_t ply);

/* --- evaluator glue (neural_network.py:128-169) --------------------------- */
void xqo_encode_board(const int8_t *board, int player,
                      float *planes /* [15][10][9] */);
void xqo_logits_to_priors(const float *logits /* [8100] */,
                          const int16_t *moves, int n, float *priors);

/* --- MCTS (self_play.py:19-175) ------------------------------------------ */
typedef void (*xqo_eval_fn)(void *ctx, int n_leaves,
                            const int8_t *boards /* [n][90] */,
                            const int32_t *players,
                            const int16_t *moves /* [n][128] */,
                            const int32_t *n_moves,
                            float *priors /* [n][128] */, double *values);

/* Returns the number of root children (0 if the root is terminal); fills
 * root_moves/root_visits in legal-move order.  stats (optional, [4]):
 * sims run, NN leaves queued, predict_batch calls, terminal backups. */
int xqo_mcts_search(const xqo_state *env, int num_simulations,
                    xqo_eval_fn eval, void *ctx, int16_t *root_moves,
                    int32_t *root_visits, int64_t *stats);

/* Built-in deterministic evaluator (hash priors, hash value) used by the MCTS
 * CPU baseline and by GPU parity tests; mirrored by the CUDA test evaluator. */
void xqo_hash_eval(void *ctx, int n_leaves, const int8_t *boards,
                   const int32_t *players, const int16_t *moves,
                   const int32_t *n_moves, float *priors, double *values);

#ifdef __cplusplus
}
#endif
#endif
