/*
 * include/xq_b200.h — C ABI of libxq_b200.so (B200 / sm_100a Xiangqi engine).
 *
 * This is the drop-in boundary for the reference's self-play hot path.  The
 * reference (hpy666666/ChineseChessAI) is pure Python with no FFI of its own;
 * each entry point below names the reference function it replaces
 * (file:line relative to the reference root).  The Python classes
 * chinesechessai_b200.chess_env.ChineseChess / self_play.MCTS bind these via
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the function name ends in
 *     `_host`.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Calls are asynchronous on that stream; `_host` calls
 *     synchronise before returning.
 *   - Return value: 0 on success, negative XQ_E_* otherwise;
 *     xq_last_error() gives a thread-local message.  No hidden global state;
 *     no CPU fallback: without a CUDA device every compute call fails.
 *   - Struct-of-arrays game state, one row per game:
 *       board    int8 [n_games][XQ_BOARD_STRIDE]  row-major r*9+c, 90 used,
 *                piece codes of config.py:66-74 (0 empty, +1..+7 red
 *                K,A,B,N,R,C,P, -1..-7 black)
 *       meta     xq_meta[n_games]                 32 B, scalars of chess_env.py:17-31,62-65
 *       pos_hist uint64[n_games][hist_cap]        position_history (chess_env.py:20,338)
 *   - A move is int16 `from*90 + to` (== the policy-logit index of
 *     neural_network.py:160); move lists are in the reference's canonical
 *     order (chess_env.py:82-88 scan order, per-piece generator order).
 */
#ifndef XQ_B200_H
#define XQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XQ_ABI_VERSION 1
#define XQ_NSQ 90
#define XQ_BOARD_STRIDE 96
#define XQ_MAX_MOVES 128     /* legal-move list capacity (observed max 67) */
#define XQ_CAND_CAP 256      /* pseudo-legal candidate capacity (bound 119 for real sets) */
#define XQ_WINNER_NONE 2     /* Python None */
#define XQ_PLANES 15
#define XQ_POLICY 8100

#define XQ_E_ARG (-1)
#define XQ_E_CUDA (-2)
#define XQ_E_NODEVICE (-3)

/* end_reason classes (chess_env.py:297,359,366,373,381,389,397,404) */
enum {
  XQ_REASON_NONE = 0,
  XQ_REASON_KING_CAPTURE = 1,
  XQ_REASON_CHECKMATE = 2,
  XQ_REASON_REPETITION = 3,
  XQ_REASON_FIFTY = 4,
  XQ_REASON_STALEMATE = 5,
  XQ_REASON_PERPETUAL_CHECK = 6,
  XQ_REASON_PERPETUAL_CHASE = 7, /* never produced: chess_env.py:674 */
  XQ_REASON_MOVE_CAP = 8
};

/* xq_meta.flags */
#define XQ_F_OVERFLOW 1u /* >XQ_MAX_MOVES legal, too many candidates / own pieces, or history full.
                          * Capacities cover every position with <= 16 pieces per side; boards
                          * with more (not chess) may raise the flag in one lane mapping and not
                          * in another - results are only specified while the flag is clear. */

typedef struct {
  int8_t player;              /* current_player +1/-1           chess_env.py:62  */
  int8_t winner;              /* 1/-1/0/XQ_WINNER_NONE          chess_env.py:64  */
  uint8_t reason;             /* XQ_REASON_*                    chess_env.py:31  */
  uint8_t done;               /* last make_move returned done (sticky)           */
  int8_t red_king;            /* cached square or -1 (None)     chess_env.py:27  */
  int8_t black_king;          /*                                chess_env.py:28  */
  uint8_t flags;              /* XQ_F_*                                          */
  uint8_t reserved;
  int32_t move_count;         /* chess_env.py:63 */
  int32_t no_capture;         /* chess_env.py:21 */
  int32_t consecutive_checks; /* chess_env.py:24 */
  int32_t hist_len;           /* len(position_history)          chess_env.py:20  */
  uint32_t check_bits;        /* last 32 check_history flags, bit0 = newest (:22) */
  int32_t check_len;          /* len(check_history)                               */
} xq_meta;                    /* 32 bytes, 16-byte aligned rows */

/* step flags byte: bit0 done, bit1 reward is a Python int, bits2-3 winner+1
 * (3 = None), bits4-7 reason */
#define XQ_STEP_DONE 1u
#define XQ_STEP_REWARD_INT 2u

typedef struct {
  int32_t plies;
  int32_t winner;
  int32_t reason;
  int32_t max_legal;
  double reward_sum;   /* sequential float64 sum of step rewards */
  uint64_t digest;     /* per-ply digest chain */
  uint64_t final_hash; /* position key of final board ‖ side to move */
} xq_playout_result;   /* 40 bytes */

/* ---- library ------------------------------------------------------------ */
int xq_abi_version(void);
const char *xq_last_error(void);
/* number of CUDA devices visible, or XQ_E_NODEVICE */
int xq_device_count(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t xq_launch_count(void);

/* ---- rules engine: chess_env.py ---------------------------------------- */
/* ChineseChess.reset  (chess_env.py:14-67) for n_games boards. */
int xq_reset(int8_t *board, xq_meta *meta, int n_games, void *stream);

/* 64-bit position key of board ‖ side byte (_get_position_hash, :497-504). */
int xq_position_hash(const int8_t *board, const xq_meta *meta, uint64_t *out,
                     int n_games, void *stream);

/* ChineseChess.get_legal_moves (chess_env.py:76-121 incl. _is_move_suicide
 * :431-464, _is_in_check :506-548, _are_kings_facing :466-495).
 * moves: int16[n_games][XQ_MAX_MOVES]; n_moves: int16[n_games].
 * in_check (optional): uint8[n_games] = _is_in_check(current_player). */
int xq_legal_moves(const int8_t *board, xq_meta *meta, int16_t *moves,
                   int16_t *n_moves, uint8_t *in_check, int n_games, void *stream);

/* _is_in_check / _are_kings_facing (chess_env.py:506-548, :466-495) on the current
 * state.  out: uint8[n_games][4] = { _is_in_check(+1), _is_in_check(-1),
 * _are_kings_facing(), 0 }, attack geometry of meta.player as in the reference. */
int xq_query_checks(const int8_t *board, const xq_meta *meta, uint8_t *out, int n_games,
                    void *stream);

/* ChineseChess.make_move (chess_env.py:253-406) for one move per game.
 * move[g] < 0 leaves game g untouched (batched loops freeze finished games;
 * the reference itself has no guard).  reward: float64[n_games]; flags:
 * uint8[n_games] (see XQ_STEP_*).  next_moves/next_n (optional) receive the
 * new side's legal list, which the terminal chain (:354,:376) computes anyway. */
int xq_step(int8_t *board, xq_meta *meta, uint64_t *pos_hist, int hist_cap,
            const int16_t *move, double *reward, uint8_t *flags,
            int16_t *next_moves, int16_t *next_n, int n_games, void *stream);

/* One ply of the random-playout game loop (self_play.py:203-256 with the search replaced by
 * the shared pick rule) in ONE launch: choose a move from the legal list that the previous call
 * (or xq_legal_moves) left in moves / n_moves, make_move it (chess_env.py:253-406), and write the
 * next position's legal list back into moves / n_moves.  Equivalent to xq_pick_moves followed by
 * xq_step with next_moves = moves; finished games and games without legal moves are frozen.
 * picked (optional): the move played, -1 for frozen games. */
int xq_step_pick(int8_t *board, xq_meta *meta, uint64_t *pos_hist, int hist_cap,
                 int16_t *moves, int16_t *n_moves, uint64_t seed,
                 uint32_t first_game_id, uint32_t ply, int capture_bias,
                 double *reward, uint8_t *flags, int16_t *picked, int n_games,
                 void *stream);

/* Counter-based uniform move pick shared with the oracle:
 * philox4x32-10(key=seed, ctr=(first_game_id+g, ply,0,0)); see DESIGN.md.
 * picked[g] = -1 when n_moves[g]==0 or the game is done. */
int xq_pick_moves(const int8_t *board, const xq_meta *meta, const int16_t *moves,
                  const int16_t *n_moves, uint64_t seed, uint32_t first_game_id,
                  uint32_t ply, int capture_bias, int16_t *picked, int n_games,
                  void *stream);

/* Diagnostics for the fused playout's per-lane kernels: while `device_buf` (3 uint64 per warp of
 * the launch, zero-filled by the caller) is registered, each warp records {first start ns, last
 * end ns, SM id}.  NULL (the default) switches it off.  Not part of the reference's surface. */
int xq_debug_playout_timing(uint64_t *device_buf);

/* Fused random playout: up to max_plies x (get_legal_moves -> pick ->
 * make_move) per game in ONE launch, state in shared memory
 * (the loop of self_play.py:203-256 with the search replaced by the pick rule).
 * The library maps boards to lanes by batch size (one warp per board below
 * 24,576 boards; two lanes per board, scheduled per SM, above - with traces
 * the one-wave lane-pair kernel from 40,960 boards; environment
 * XQ_PLAYOUT_MODE = warp | tpb | pair | pairq | pairs forces one) - results
 * are identical in every mapping.
 * A game whose meta.flags gets XQ_F_OVERFLOW (more than XQ_MAX_MOVES candidate
 * moves or a full history; unreachable from legal chess positions) keeps
 * running on in-range but unspecified moves.
 * Optional per-ply traces (NULL to skip), each [n_games][max_plies]:
 *   tr_moves int16[..][XQ_MAX_MOVES], tr_n int16, tr_pick int16,
 *   tr_reward float64, tr_flags uint8, tr_boards int8[..][XQ_NSQ]. */
int xq_playout(int8_t *board, xq_meta *meta, uint64_t *pos_hist, int hist_cap,
               uint64_t seed, uint32_t first_game_id, int max_plies,
               int capture_bias, xq_playout_result *results, int16_t *tr_moves,
               int16_t *tr_n, int16_t *tr_pick, double *tr_reward,
               uint8_t *tr_flags, int8_t *tr_boards, int n_games, void *stream);

/* Same, end to end from HOST buffers: H2D of board/meta, playout, D2H of the
 * final board/meta/results.  pos_hist is device scratch owned by the call. */
int xq_playout_host(int8_t *board_h, xq_meta *meta_h, uint64_t seed,
                    uint32_t first_game_id, int max_plies, int capture_bias,
                    xq_playout_result *results_h, int n_games, int device);

/* ---- evaluator glue: neural_network.py ---------------------------------- */
/* ChessNet.encode_board (neural_network.py:128-146): planes
 * float32[n][15][10][9]; or bf16 (as uint16) when out_bf16 != 0. */
int xq_encode_planes(const int8_t *board, int board_stride, const int8_t *player,
                     int player_stride, void *planes, int out_bf16, int n,
                     void *stream);

/* Same planes for the inference copy of the network: bf16, channels-last
 * [n][10][9][16] (the 15 planes + one zero channel), i.e. the memory a torch
 * tensor of shape [n,16,10,9] in channels_last format has - 32 contiguous
 * bytes per square, the layout the tensor-core convolution reads directly. */
int xq_encode_planes_nhwc16(const int8_t *board, int board_stride, const int8_t *player,
                            int player_stride, void *planes, int n, void *stream);

/* ChessNet._logits_to_move_probs (neural_network.py:148-169): gather the
 * legal moves' logits and softmax in float32.  logits: float32 or bf16 rows of
 * logits_stride >= XQ_POLICY elements (a padded policy head may be passed as is);
 * priors float32[n][XQ_MAX_MOVES]. */
int xq_policy_priors(const void *logits, int logits_bf16, int logits_stride,
                     const int16_t *moves, int moves_stride, const int16_t *n_moves,
                     float *priors, int n, void *stream);

/* Residual-block epilogue of ChessNet (neural_network.py:181-187) for the bf16 channels-last
 * inference copy: out = relu(y + bias[c] + x) in one pass (y = conv2 output without bias,
 * x = block input, c = innermost index mod channels).  cuDNN's fused conv+add+ReLU runs the same
 * convolution at half the tensor-pipe utilisation (profiles/r1/nn_kernels_tensor_pipe_cfg3.csv);
 * plain conv + this HBM-bound pass is faster.  n_elems must be a multiple of 8, channels too. */
int xq_bias_residual_relu_bf16(const void *y, const void *x, const void *bias, void *out,
                               int64_t n_elems, int channels, void *stream);

/* encode_board (neural_network.py:128-146) + the network's first layer conv1 -> bn1 -> ReLU
 * (neural_network.py:25-27,:60) in ONE kernel for the bf16 inference copy.  The 15 input planes
 * are one-hot, so the 3x3 convolution at a square is a SUM of weight columns selected by the
 * pieces on its 9 neighbours: no planes tensor and no GEMM; float32 accumulation of the same
 * bf16 weights cuDNN would use.  (A tensor-core form of the same idea, mma.sync m16n8k16 with
 * the one-hot operand built in registers, was measured slower: DESIGN.md section 7.)
 * table: bf16 [9 taps][16 input channels][channels] with tap = kh*3+kw of the BN-folded
 * 3x3/stride 1/pad 1 weight (input channel 15 unused); bias: float32 [2][90][channels] =
 * the folded bias for black to move and, for red to move ([1]), the bias plus the
 * side-to-move plane's weights summed over the taps that lie on the board at that square;
 * out: bf16 [n][10][9][channels] (channels-last).  channels == 128 (config.py:33). */
int xq_stem_lookup_bf16(const int8_t *board, int board_stride, const int8_t *player,
                        int player_stride, const void *table, const float *bias, void *out,
                        int channels, int n, void *stream);

/* ---- MCTS: self_play.py:19-175 ------------------------------------------- */
/* One flat node pool per game ("tree"), caller-allocated device memory of
 * xq_mcts_tree_bytes(num_simulations) bytes per game (16-byte aligned; every
 * call takes the same num_simulations, which fixes the per-game stride):
 * header, nodes (MCTSNode :19-28: parent, move, children contiguous in
 * legal-move order, visit_count int32, value_sum float64, prior_prob float32)
 * and the env state of every expanded node (so a simulation steps the rules
 * engine once at its leaf instead of replaying the path from the root as
 * MCTS.search :115-119 does; results are identical because make_move is
 * deterministic).  Wave semantics follow :101-148 exactly: no virtual loss,
 * terminal leaves backed up immediately in sim order, the (single) network
 * leaf of a wave expanded once and backed up once per simulation that
 * reached it, float32 PUCT in NumPy>=2 op order with first-max tie-break
 * (:40-59), float64 value_sum. */
int64_t xq_mcts_tree_bytes(int num_simulations);

/* root = MCTS._copy_env(env) (self_play.py:156-175): board, side, move_count,
 * winner, king caches, no_capture copied; histories empty, consecutive_checks 0.
 * active (optional uint8[n_games]): 0 = skip this game (search returns no children). */
int xq_mcts_init(void *trees, int num_simulations, const int8_t *board,
                 const xq_meta *meta, const uint8_t *active, int n_games, void *stream);

/* One wave (self_play.py:103-139): runs min(wave_size, remaining) simulations
 * per game up to and including the first non-terminal leaf.  Outputs that leaf
 * for the evaluator: leaf_board int8[n][XQ_BOARD_STRIDE], leaf_player int8[n],
 * leaf_moves int16[n][XQ_MAX_MOVES], leaf_n int16[n] (0 = nothing to evaluate),
 * leaf_mult int16[n] = simulations of this wave that reached the leaf. */
int xq_mcts_select(void *trees, int num_simulations, int wave_size, int8_t *leaf_board, int8_t *leaf_player,
                   int16_t *leaf_moves, int16_t *leaf_n, int16_t *leaf_mult, int n_games,
                   void *stream);

/* Expand + backup (self_play.py:142-148, :61-80).  priors float32[n][XQ_MAX_MOVES]
 * in leaf_moves order; values: float64 (values_f32 == 0) or float32, one per game
 * (values_per_game == 1) or one per queued simulation (values_per_game == 8, the
 * reference passes the duplicated leaf to predict_batch once per simulation). */
int xq_mcts_backup(void *trees, int num_simulations, const int16_t *leaf_moves, const int16_t *leaf_n,
                   const float *priors, const void *values, int values_f32,
                   int values_per_game, int n_games, void *stream);

/* Leaf compaction for ragged batches.  The reference only sends simulations that reached a
 * non-terminal leaf to the network (self_play.py:126-139: terminal leaves are backed up on the
 * spot); in a batch, finished games and waves that ended on terminal leaves have leaf_n == 0.
 * xq_compact_leaves lists the games with leaf_n > 0 in game order (idx int32[n_games], *count)
 * and gives every game its row in that list (row_of_game int32[n_games], -1 if none);
 * xq_gather_leaves copies those rows of the four leaf arrays into compact arrays of `rows` rows
 * (rows >= *count; padding rows are empty positions); the evaluator then runs on `rows` rows and
 * xq_mcts_backup_rows = xq_mcts_backup reading priors float32[rows][XQ_MAX_MOVES] / values[rows]
 * (one per game) through row_of_game.  Visit counts are identical to the uncompacted search. */
int xq_compact_leaves(const int16_t *leaf_n, int n_games, int32_t *idx, int32_t *row_of_game,
                      int32_t *count, void *stream);
int xq_gather_leaves(const int32_t *idx, const int32_t *count, int rows, const int8_t *leaf_board,
                     const int8_t *leaf_player, const int16_t *leaf_moves, const int16_t *leaf_n,
                     int8_t *out_board, int8_t *out_player, int16_t *out_moves, int16_t *out_n,
                     void *stream);
int xq_mcts_backup_rows(void *trees, int num_simulations, const int16_t *leaf_moves,
                        const int16_t *leaf_n, const float *priors, const void *values,
                        int values_f32, const int32_t *row_of_game, int n_games, void *stream);

/* {move: child.visit_count} of the root (self_play.py:151-154), legal-move order. */
int xq_mcts_root_visits(const void *trees, int num_simulations, int16_t *moves, int32_t *visits,
                        int16_t *n_children, int n_games, void *stream);

/* Move choice of self_play_game (self_play.py:219-243) for a batch: p_i = visits_i^(1/T) /
 * sum (float64), or a one-hot on the first maximum when T < 0.01 (:224-227); the index is drawn
 * by inverse CDF with u = philox4x32-10(key=seed, ctr=(first_game_id+g, ply, 1, 0)) in [0,1),
 * so a game's trajectory depends on (seed, game id) only — not on which GPU or batch plays it.
 * (The reference draws from the process-global np.random stream; parity is defined on the
 * visit counts)  chosen[g] = index into the root's move list, or -1 when the game
 * is inactive or has no children.  active: optional uint8[n_games]. */
int xq_sample_moves(const int32_t *visits, const int16_t *n_children, const uint8_t *active,
                    double temperature, uint64_t seed, uint32_t first_game_id, uint32_t ply,
                    int16_t *chosen, int n_games, void *stream);

/* The rest of one ply of the batched game loop (self_play.py:203-256), so that a ply is
 * search + 4 launches instead of ~25 framework launches:
 * xq_selfplay_commit records what the reference appends to game_data (:229-231) — the board
 * before the move, the side to move, the root's move list and visit counts — into the caller's
 * per-ply rows, resolves chosen[g] to the move to play (move[g] = -1 for a game that is
 * inactive or has no legal move) and clears *any_active;
 * xq_selfplay_finish, called after xq_step(move), retires games: active[g] stays 1 only if
 * the game was stepped and the step did not end it (:254-255), and *any_active receives the
 * NUMBER of games still running (0 = the batch is over; the count also bounds the rows the
 * next searches have to evaluate, see xq_compact_leaves).
 * rec_board int8[n][90], rec_player int8[n], rec_moves int16[n][XQ_MAX_MOVES],
 * rec_visits int32[n][XQ_MAX_MOVES], rec_n int16[n], rec_played uint8[n], rec_move int16[n]. */
int xq_selfplay_commit(const int16_t *root_moves, const int32_t *root_visits,
                       const int16_t *root_n, const int16_t *chosen, const int8_t *board,
                       const xq_meta *meta, int8_t *rec_board, int8_t *rec_player,
                       int16_t *rec_moves, int32_t *rec_visits, int16_t *rec_n,
                       uint8_t *rec_played, int16_t *rec_move, int16_t *move,
                       int32_t *any_active, int n_games, void *stream);
int xq_selfplay_finish(const int16_t *move, const uint8_t *step_flags, uint8_t *active,
                       int32_t *any_active, int n_games, void *stream);

/* Deterministic test evaluator (hashed or flat priors, hashed value) identical
 * to the oracle's, so MCTS parity can be checked at batch scale without a net. */
int xq_hash_eval(const int8_t *board, int board_stride, const int8_t *player,
                 const int16_t *moves, const int16_t *n_moves, int flat, float *priors,
                 double *values, int n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XQ_B200_H */
