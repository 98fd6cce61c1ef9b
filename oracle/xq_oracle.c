/*
 * oracle/xq_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Literal plain-C restatement of the reference algorithms on the hot path:
 *   rules  : /root/reference/chess_env.py      (ChineseChess)
 *   search : /root/reference/self_play.py:19-175 (MCTSNode, MCTS)
 *   glue   : /root/reference/neural_network.py:128-169 (encode, priors)
 * It deliberately keeps the reference's *forward* formulation (regenerate every
 * opposing piece's pseudo-moves, simulate each candidate on a board copy) so
 * that it is independent of the CUDA kernels, which use an inverse
 * (king-outward) attack test.  Dead work whose result is never read
 * (_get_threatened_pieces/_is_protected/chase_history, chess_env.py:262,344,
 * 550-596,664-681; _get_material_advantage :739-768) is not reproduced.
 *
 * Parity: PINNED against the imported reference — see oracle/xq_oracle.h.
 */
#include "xq_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define R_KING 1
#define R_ADVISOR 2
#define R_BISHOP 3
#define R_KNIGHT 4
#define R_ROOK 5
#define R_CANNON 6
#define R_PAWN 7

typedef struct {
  int r, c;
} rc_t;

static inline int on_board(int r, int c) {
  return r >= 0 && r < XQO_ROWS && c >= 0 && c < XQO_COLS;
}
static inline int at(const xqo_state *s, int r, int c) {
  return s->board[r * XQO_COLS + c];
}

/* ------------------------------------------------------------------------ */
/* 64-bit position key.  The reference uses Python's salted hash() of
 * board.tobytes()+player byte (chess_env.py:497-504); only equality is
 * observable, so any injective-in-practice 64-bit key is equivalent.  This
 * exact function is part of the repo's spec so that histories can
 * be compared bit-for-bit between oracle and device. */
static inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

uint64_t xqo_position_hash(const int8_t *board, int player) {
  uint64_t h = mix64(0x7000ULL + (player == 1 ? 0u : 1u)); /* :503 */
  for (int sq = 0; sq < XQO_NSQ; ++sq) {
    int p = board[sq];
    if (p != 0) h ^= mix64((uint64_t)((p + 8) * 128 + sq));
  }
  return h;
}

/* ------------------------------------------------------------------------ */
/* chess_env.py:14-67 */
void xqo_reset(xqo_state *s) {
  memset(s, 0, sizeof(*s));
  static const int8_t back[9] = {R_ROOK,    R_KNIGHT, R_BISHOP, R_ADVISOR, R_KING,
                                 R_ADVISOR, R_BISHOP, R_KNIGHT, R_ROOK};
  for (int c = 0; c < 9; ++c) {
    s->board[9 * 9 + c] = back[c];
    s->board[0 * 9 + c] = (int8_t)-back[c];
  }
  s->board[7 * 9 + 1] = s->board[7 * 9 + 7] = R_CANNON;
  s->board[2 * 9 + 1] = s->board[2 * 9 + 7] = -R_CANNON;
  for (int c = 0; c < 9; c += 2) {
    s->board[6 * 9 + c] = R_PAWN;
    s->board[3 * 9 + c] = -R_PAWN;
  }
  s->red_king = 9 * 9 + 4;
  s->black_king = 0 * 9 + 4;
  s->player = 1;
  s->move_count = 0;
  s->winner = XQO_WINNER_NONE;
  s->reason = XQO_REASON_NONE;
}

/* ---- pseudo-move generators; geometry follows s->player, NOT the piece's
 * colour (chess_env.py:127,143,159,240).  Targets may be off-board. -------- */

/* chess_env.py:123-138 */
static int king_moves(const xqo_state *s, int r, int c, rc_t *out) {
  static const int d[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};
  int lo = s->player == 1 ? 7 : 0, hi = s->player == 1 ? 10 : 3, n = 0;
  for (int i = 0; i < 4; ++i) {
    int nr = r + d[i][0], nc = c + d[i][1];
    if (lo <= nr && nr < hi && 3 <= nc && nc < 6) out[n++] = (rc_t){nr, nc};
  }
  return n;
}

/* chess_env.py:140-154 */
static int advisor_moves(const xqo_state *s, int r, int c, rc_t *out) {
  static const int d[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
  int lo = s->player == 1 ? 7 : 0, hi = s->player == 1 ? 10 : 3, n = 0;
  for (int i = 0; i < 4; ++i) {
    int nr = r + d[i][0], nc = c + d[i][1];
    if (lo <= nr && nr < hi && 3 <= nc && nc < 6) out[n++] = (rc_t){nr, nc};
  }
  return n;
}

/* chess_env.py:156-176 */
static int bishop_moves(const xqo_state *s, int r, int c, rc_t *out) {
  static const int d[4][2] = {{2, 2}, {2, -2}, {-2, 2}, {-2, -2}};
  int river = s->player == 1 ? 5 : 4, n = 0;
  for (int i = 0; i < 4; ++i) {
    int nr = r + d[i][0], nc = c + d[i][1];
    if (!on_board(nr, nc)) continue;
    if (s->player == 1 && nr < river) continue;
    if (s->player == -1 && nr >= river) continue;
    if (at(s, r + d[i][0] / 2, c + d[i][1] / 2) == 0) out[n++] = (rc_t){nr, nc};
  }
  return n;
}

/* chess_env.py:178-197 */
static int knight_moves(const xqo_state *s, int r, int c, rc_t *out) {
  static const int d[8][4] = {{2, 1, 1, 0},   {2, -1, 1, 0}, {-2, 1, -1, 0},
                              {-2, -1, -1, 0}, {1, 2, 0, 1},  {-1, 2, 0, 1},
                              {1, -2, 0, -1},  {-1, -2, 0, -1}};
  int n = 0;
  for (int i = 0; i < 8; ++i) {
    int br = r + d[i][2], bc = c + d[i][3];
    if (on_board(br, bc) && at(s, br, bc) == 0)
      out[n++] = (rc_t){r + d[i][0], c + d[i][1]};
  }
  return n;
}

static const int RAY[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};

/* chess_env.py:199-213 */
static int rook_moves(const xqo_state *s, int r, int c, rc_t *out) {
  int n = 0;
  for (int i = 0; i < 4; ++i)
    for (int step = 1; step < 10; ++step) {
      int nr = r + RAY[i][0] * step, nc = c + RAY[i][1] * step;
      if (!on_board(nr, nc)) break;
      out[n++] = (rc_t){nr, nc};
      if (at(s, nr, nc) != 0) break;
    }
  return n;
}

/* chess_env.py:215-235 */
static int cannon_moves(const xqo_state *s, int r, int c, rc_t *out) {
  int n = 0;
  for (int i = 0; i < 4; ++i) {
    int jumped = 0;
    for (int step = 1; step < 10; ++step) {
      int nr = r + RAY[i][0] * step, nc = c + RAY[i][1] * step;
      if (!on_board(nr, nc)) break;
      if (at(s, nr, nc) == 0) {
        if (!jumped) out[n++] = (rc_t){nr, nc};
      } else if (!jumped) {
        jumped = 1;
      } else {
        out[n++] = (rc_t){nr, nc};
        break;
      }
    }
  }
  return n;
}

/* chess_env.py:237-251 */
static int pawn_moves(const xqo_state *s, int r, int c, rc_t *out) {
  int n = 0;
  if (s->player == 1) {
    out[n++] = (rc_t){r - 1, c};
    if (r < 5) {
      out[n++] = (rc_t){r, c - 1};
      out[n++] = (rc_t){r, c + 1};
    }
  } else {
    out[n++] = (rc_t){r + 1, c};
    if (r >= 5) {
      out[n++] = (rc_t){r, c - 1};
      out[n++] = (rc_t){r, c + 1};
    }
  }
  return n;
}

/* dispatch of chess_env.py:95-108 and :527-542 */
static int pseudo_moves(const xqo_state *s, int r, int c, int piece, rc_t *out) {
  switch (abs(piece)) {
    case R_KING: return king_moves(s, r, c, out);
    case R_ADVISOR: return advisor_moves(s, r, c, out);
    case R_BISHOP: return bishop_moves(s, r, c, out);
    case R_KNIGHT: return knight_moves(s, r, c, out);
    case R_ROOK: return rook_moves(s, r, c, out);
    case R_CANNON: return cannon_moves(s, r, c, out);
    case R_PAWN: return pawn_moves(s, r, c, out);
    default: return 0;
  }
}

/* chess_env.py:506-548 */
int xqo_is_in_check(const xqo_state *s, int player) {
  int king = player == 1 ? s->red_king : s->black_king;
  if (king < 0) return 0;
  int kr = king / 9, kc = king % 9;
  rc_t mv[40];
  for (int r = 0; r < XQO_ROWS; ++r)
    for (int c = 0; c < XQO_COLS; ++c) {
      int piece = at(s, r, c);
      if (piece * player < 0) {
        int n = pseudo_moves(s, r, c, piece, mv);
        for (int i = 0; i < n; ++i)
          if (mv[i].r == kr && mv[i].c == kc) return 1;
      }
    }
  return 0;
}

/* chess_env.py:466-495 */
int xqo_kings_facing(const xqo_state *s) {
  if (s->red_king < 0 || s->black_king < 0) return 0;
  int rr = s->red_king / 9, rc = s->red_king % 9;
  int br = s->black_king / 9, bc = s->black_king % 9;
  if (rc != bc) return 0;
  int lo = rr < br ? rr : br, hi = rr < br ? br : rr;
  for (int r = lo + 1; r < hi; ++r)
    if (at(s, r, rc) != 0) return 0;
  return 1;
}

/* chess_env.py:431-464 */
static int is_move_suicide(xqo_state *s, int fr, int fc, int tr, int tc) {
  int8_t backup[XQO_NSQ];
  memcpy(backup, s->board, XQO_NSQ);
  int b_red = s->red_king, b_black = s->black_king;
  int moving = at(s, fr, fc);
  s->board[tr * 9 + tc] = (int8_t)moving;
  s->board[fr * 9 + fc] = 0;
  if (moving == R_KING)
    s->red_king = tr * 9 + tc;
  else if (moving == -R_KING)
    s->black_king = tr * 9 + tc;
  int in_check = xqo_is_in_check(s, s->player);
  int facing = xqo_kings_facing(s);
  memcpy(s->board, backup, XQO_NSQ);
  s->red_king = b_red;
  s->black_king = b_black;
  return in_check || facing;
}

/* chess_env.py:76-121; with_filter=0 counts candidates that pass :113,:116 */
static int gen_moves(xqo_state *s, int16_t *moves, int with_filter) {
  int n = 0;
  rc_t mv[40];
  for (int r = 0; r < XQO_ROWS; ++r)
    for (int c = 0; c < XQO_COLS; ++c) {
      int piece = at(s, r, c);
      if (piece * s->player > 0) {
        int k = pseudo_moves(s, r, c, piece, mv);
        for (int i = 0; i < k; ++i) {
          int tr = mv[i].r, tc = mv[i].c;
          if (!on_board(tr, tc)) continue;
          if (at(s, tr, tc) * s->player > 0) continue;
          if (with_filter && is_move_suicide(s, r, c, tr, tc)) continue;
          if (n < XQO_MAX_MOVES && moves)
            moves[n] = (int16_t)((r * 9 + c) * 90 + tr * 9 + tc);
          else if (moves)
            s->overflow = 1;
          ++n;
        }
      }
    }
  return n;
}

int xqo_legal_moves(xqo_state *s, int16_t *moves) {
  int n = gen_moves(s, moves, 1);
  return n > XQO_MAX_MOVES ? XQO_MAX_MOVES : n;
}
int xqo_pseudo_count(xqo_state *s) { return gen_moves(s, NULL, 0); }

/* chess_env.py:683-737; evaluated before the side switch */
static double position_change(const xqo_state *s, int fr, int fc, int tr, int tc) {
  int type = abs(at(s, tr, tc));
  double score = 0;
  int advance = s->player == 1 ? fr - tr : tr - fr;
  if (advance > 0) {
    if (type == R_PAWN)
      score += advance * 2.0;
    else if (type == R_ROOK || type == R_CANNON)
      score += advance * 1.5;
    else if (type == R_KNIGHT)
      score += advance * 1.0;
  }
  if (tc >= 3 && tc <= 5) {
    score += 1.5;
    if (3 <= tr && tr <= 6) score += 1.0;
  }
  if (type == R_PAWN) {
    if (s->player == 1 && tr < 5)
      score += 3.0;
    else if (s->player == -1 && tr >= 5)
      score += 3.0;
  }
  int ok = s->player == 1 ? s->black_king : s->red_king;
  if (ok >= 0) {
    int kr = ok / 9, kc = ok % 9;
    int old_d = abs(fr - kr) + abs(fc - kc);
    int new_d = abs(tr - kr) + abs(tc - kc);
    if (new_d < old_d) score += (old_d - new_d) * 0.5;
  }
  return score;
}

/* chess_env.py:646-662 */
static int perpetual_check(const xqo_state *s) {
  if (s->check_len < 12) return 0;
  int cnt = 0;
  for (int i = s->check_len - 12; i < s->check_len; ++i) cnt += s->check_hist[i] != 0;
  return cnt >= 10;
}

/* chess_env.py:253-406 */
void xqo_make_move(xqo_state *s, int move, xqo_step_result *out) {
  int from = move / 90, to = move % 90;
  int fr = from / 9, fc = from % 9, tr = to / 9, tc = to % 9;

  int captured = s->board[to];           /* :265 */
  int moving = s->board[from];           /* :266 */
  s->board[to] = (int8_t)moving;
  s->board[from] = 0;

  if (moving == R_KING) /* :271-279 */
    s->red_king = to;
  else if (moving == -R_KING)
    s->black_king = to;
  if (captured == R_KING)
    s->red_king = -1;
  else if (captured == -R_KING)
    s->black_king = -1;

  if (captured != 0) /* :282-285 */
    s->no_capture = 0;
  else
    s->no_capture += 1;

  double reward = 0;
  int is_int = 1, done = 0;

  if (abs(captured) == R_KING) { /* :292-297 */
    s->winner = s->player;
    reward = 100;
    done = 1;
    s->reason = XQO_REASON_KING_CAPTURE;
  } else if (captured != 0) { /* :300-314 */
    double base = 0;
    switch (abs(captured)) {
      case R_ROOK: base = 9; break;
      case R_CANNON: base = 4.5; break;
      case R_KNIGHT: base = 4; break;
      case R_BISHOP: base = 2; break;
      case R_ADVISOR: base = 2; break;
      case R_PAWN: base = 1; break;
    }
    reward = base * 2.0;
    is_int = 0;
    if (abs(captured) == R_ADVISOR || abs(captured) == R_BISHOP) reward += 3.0;
  }

  int is_checking = xqo_is_in_check(s, -s->player); /* :317 */
  if (!done && is_checking) {                       /* :318-327 */
    if (s->consecutive_checks == 0) {
      reward += 15.0;
      is_int = 0;
    } else if (s->consecutive_checks == 1) {
      reward += 10.0;
      is_int = 0;
    } else if (s->consecutive_checks == 2) {
      reward += 5.0;
      is_int = 0;
    }
    s->consecutive_checks += 1;
  } else { /* :328-335 */
    s->consecutive_checks = 0;
    if (captured == 0 && !done) {
      reward += position_change(s, fr, fc, tr, tc) * 0.01;
      is_int = 0;
    }
  }

  /* :338 — hashed with the MOVER's player byte (before the switch) */
  if (s->pos_len < XQO_HIST_CAP)
    s->pos_hist[s->pos_len++] = xqo_position_hash(s->board, s->player);
  else
    s->overflow = 1;
  if (s->check_len < XQO_HIST_CAP) /* :341 */
    s->check_hist[s->check_len++] = (uint8_t)is_checking;
  else
    s->overflow = 1;

  s->player = -s->player; /* :348-349 */
  s->move_count += 1;

  if (!done) { /* :352-397 */
    int16_t mv[XQO_MAX_MOVES];
    int n_legal = gen_moves(s, mv, 1);
    int in_check_now = -1;
    if (n_legal == 0) in_check_now = xqo_is_in_check(s, s->player);
    if (n_legal == 0 && in_check_now) { /* :354 / :614-628 */
      done = 1;
      reward = 200;
      is_int = 1;
      s->winner = -s->player;
      s->reason = XQO_REASON_CHECKMATE;
    } else {
      uint64_t h = xqo_position_hash(s->board, s->player); /* :603 */
      int cnt = 0;
      for (int i = 0; i < s->pos_len; ++i) cnt += s->pos_hist[i] == h;
      if (cnt >= 3) { /* :362 */
        done = 1;
        reward = 0;
        is_int = 1;
        s->winner = 0;
        s->reason = XQO_REASON_REPETITION;
      } else if (s->no_capture >= 100) { /* :369 / :612 */
        done = 1;
        reward = 0;
        is_int = 1;
        s->winner = 0;
        s->reason = XQO_REASON_FIFTY;
      } else if (n_legal == 0 && !in_check_now) { /* :376 / :630-644 */
        done = 1;
        reward = 100;
        is_int = 1;
        s->winner = -s->player;
        s->reason = XQO_REASON_STALEMATE;
      } else if (perpetual_check(s)) { /* :384 */
        done = 1;
        reward = -10;
        is_int = 1;
        s->winner = -s->player;
        s->reason = XQO_REASON_PERPETUAL_CHECK;
      } /* :392 perpetual chase: always False (:674) */
    }
  }

  if (!done && s->move_count >= 70) { /* :400-404 */
    done = 1;
    reward = -2;
    is_int = 1;
    s->winner = 0;
    s->reason = XQO_REASON_MOVE_CAP;
  }

  out->reward = reward;
  out->reward_is_int = is_int;
  out->done = done;
}

/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) — counter-based pick shared with the
 * CUDA playout kernel (SURVEY.md §8d cfg 1). */
void xqo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

int xqo_pick_move(const xqo_state *s, const int16_t *moves, int n,
                  uint64_t seed, uint32_t game_id, uint32_t ply,
                  int capture_bias) {
  uint32_t x[4];
  xqo_philox4x32(game_id, ply, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  if (capture_bias > 0 && (int)(x[1] & 0xFFu) < capture_bias) {
    int ncap = 0;
    for (int i = 0; i < n; ++i) ncap += s->board[moves[i] % 90] != 0;
    if (ncap > 0) {
      int k = (int)(x[0] % (uint32_t)ncap);
      for (int i = 0; i < n; ++i)
        if (s->board[moves[i] % 90] != 0 && k-- == 0) return i;
    }
  }
  return (int)(x[0] % (uint32_t)n);
}

/* self_play.py:219-243: counts ** (1/T) / sum, np.random.choice by inverse CDF
 * (searchsorted(cumsum(p), u, side="right")); u from the shared counter-based generator. */
int xqo_sample_move(const int32_t *visits, int n, double temperature, uint64_t seed,
                    uint32_t game_id, uint32_t ply) {
  uint32_t x[4];
  xqo_philox4x32(game_id, ply, 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  if (n <= 0) return -1;
  if (temperature < 0.01) {
    int pick = 0;
    for (int i = 1; i < n; ++i)
      if (visits[i] > visits[pick]) pick = i;
    return pick;
  }
  double inv_t = 1.0 / temperature, total = 0.0;
  for (int i = 0; i < n; ++i) total += pow((double)visits[i], inv_t);
  if (!(total > 0.0)) return (int)(x[0] % (uint32_t)n);
  double u = (double)((((uint64_t)x[0] << 32) | x[1]) >> 11) * (1.0 / 9007199254740992.0);
  double cdf = 0.0;
  for (int i = 0; i < n; ++i) {
    cdf += pow((double)visits[i], inv_t) / total;
    if (cdf > u) return i;
  }
  return n - 1;
}

static inline uint64_t dbits(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}

/* Digest of one ply: position-weighted 32-bit sum of the ordered move
 * list, packed with n / picked move / ply, then combined with the reward bits, the outcome
 * flags and the key of the resulting position by odd 64-bit multipliers.  The chain is
 * digest = mix64(digest ^ word). */
static uint64_t ply_digest(int ply, const int16_t *moves, int n, int pick_move,
                           const xqo_state *after, const xqo_step_result *r) {
  uint32_t lsum = 0;
  for (int i = 0; i < n; ++i) lsum += (uint32_t)(moves[i] + 1) * (uint32_t)(2 * i + 1);
  uint64_t a = (uint64_t)lsum | ((uint64_t)n << 32) | ((uint64_t)pick_move << 40) |
               ((uint64_t)(ply + 1) << 54);
  uint64_t c = (uint64_t)(r->done & 1) | ((uint64_t)(after->winner + 2) << 8) |
               ((uint64_t)after->reason << 16) | ((uint64_t)(r->reward_is_int & 1) << 24);
  return a * 0x9E3779B97F4A7C15ULL + dbits(r->reward) * 0xC2B2AE3D27D4EB4FULL +
         c * 0x165667B19E3779F9ULL +
         xqo_position_hash(after->board, after->player) * 0x27D4EB2F165667C5ULL;
}

void xqo_playout(xqo_state *s, uint64_t seed, uint32_t game_id, int max_plies,
                 int capture_bias, xqo_playout_result *res, int16_t *trace_moves,
                 int16_t *trace_n, int16_t *trace_pick, double *trace_reward,
                 uint8_t *trace_flags, int8_t *trace_boards) {
  int16_t moves[XQO_MAX_MOVES];
  memset(res, 0, sizeof(*res));
  int ply = 0;
  for (; ply < max_plies; ++ply) {
    int n = xqo_legal_moves(s, moves);
    if (n == 0) break; /* self_play.py:207 */
    if (n > res->max_legal) res->max_legal = n;
    int idx = xqo_pick_move(s, moves, n, seed, game_id, (uint32_t)ply, capture_bias);
    xqo_step_result r;
    xqo_make_move(s, moves[idx], &r);
    res->reward_sum += r.reward;
    res->digest = mix64(res->digest ^ ply_digest(ply, moves, n, moves[idx], s, &r));
    if (trace_moves) memcpy(trace_moves + (size_t)ply * XQO_MAX_MOVES, moves, n * sizeof(int16_t));
    if (trace_n) trace_n[ply] = (int16_t)n;
    if (trace_pick) trace_pick[ply] = moves[idx];
    if (trace_reward) trace_reward[ply] = r.reward;
    if (trace_flags)
      trace_flags[ply] = (uint8_t)((r.done & 1) | ((r.reward_is_int & 1) << 1) |
                                   ((s->winner + 1) << 2) | (s->reason << 4));
    if (trace_boards) memcpy(trace_boards + (size_t)ply * XQO_NSQ, s->board, XQO_NSQ);
    if (r.done) {
      ++ply;
      break;
    }
  }
  res->plies = ply;
  res->winner = s->winner;
  res->reason = s->reason;
  res->final_hash = xqo_position_hash(s->board, s->player);
}

typedef struct {
  int begin, end;
  uint32_t first_id;
  uint64_t seed;
  int max_plies, bias;
  xqo_playout_result *res;
  int64_t plies;
} job_t;

static void *playout_worker(void *arg) {
  job_t *j = (job_t *)arg;
  xqo_state *s = (xqo_state *)malloc(sizeof(xqo_state));
  for (int g = j->begin; g < j->end; ++g) {
    xqo_reset(s);
    xqo_playout(s, j->seed, j->first_id + (uint32_t)g, j->max_plies, j->bias,
                &j->res[g], NULL, NULL, NULL, NULL, NULL, NULL);
    j->plies += j->res[g].plies;
  }
  free(s);
  return NULL;
}

int64_t xqo_playout_many(int n_games, uint32_t first_game_id, uint64_t seed,
                         int max_plies, int capture_bias, int n_threads,
                         xqo_playout_result *results) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < n_threads; ++t) {
    jobs[t] = (job_t){(int)((int64_t)n_games * t / n_threads),
                      (int)((int64_t)n_games * (t + 1) / n_threads),
                      first_game_id, seed, max_plies, capture_bias, results, 0};
    pthread_create(&th[t], NULL, playout_worker, &jobs[t]);
  }
  int64_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    total += jobs[t].plies;
  }
  return total;
}

/* ------------------------------------------------------------------------ */
/* neural_network.py:128-146 — no perspective flip; plane 14 = (player==1) */
void xqo_encode_board(const int8_t *board, int player, float *planes) {
  for (int i = 1; i <= 7; ++i)
    for (int sq = 0; sq < XQO_NSQ; ++sq) {
      planes[(i - 1) * XQO_NSQ + sq] = board[sq] == i ? 1.0f : 0.0f;
      planes[(i + 6) * XQO_NSQ + sq] = board[sq] == -i ? 1.0f : 0.0f;
    }
  for (int sq = 0; sq < XQO_NSQ; ++sq) planes[14 * XQO_NSQ + sq] = player == 1 ? 1.0f : 0.0f;
}

/* neural_network.py:148-169 — gather by packed index (== from*90+to, :160),
 * float32 max-subtracted softmax.  numpy's SIMD exp / pairwise sum are not
 * bit-reproducible; tests compare with 1e-6 relative tolerance (B.5). */
void xqo_logits_to_priors(const float *logits, const int16_t *moves, int n, float *priors) {
  if (n == 0) return;
  float mx = -INFINITY;
  for (int i = 0; i < n; ++i) {
    priors[i] = logits[moves[i]];
    if (priors[i] > mx) mx = priors[i];
  }
  float sum = 0.0f;
  for (int i = 0; i < n; ++i) {
    priors[i] = expf(priors[i] - mx);
    sum += priors[i];
  }
  for (int i = 0; i < n; ++i) priors[i] = priors[i] / sum;
}

/* ------------------------------------------------------------------------ */
/* MCTS — self_play.py:19-175, literal (one env copy + full replay per sim). */
typedef struct {
  int parent;       /* :22 */
  int move;         /* :23 */
  int first_child;  /* children dict (:24), contiguous, insertion order */
  int n_children;
  int visits;       /* :26 */
  double value_sum; /* :27 */
  float prior;      /* :28 — numpy.float32 from _logits_to_move_probs */
} node_t;

typedef struct {
  node_t *a;
  int n, cap;
} pool_t;

static int pool_new(pool_t *p, int parent, int move, float prior) {
  if (p->n == p->cap) {
    p->cap = p->cap ? p->cap * 2 : 1024;
    p->a = (node_t *)realloc(p->a, sizeof(node_t) * (size_t)p->cap);
  }
  p->a[p->n] = (node_t){parent, move, -1, 0, 0, 0.0, prior};
  return p->n++;
}

/* self_play.py:40-59; NumPy>=2 promotion: every PUCT op is float32 (B.3) */
static int select_child(const pool_t *p, int node) {
  const node_t *nd = &p->a[node];
  float best = -INFINITY;
  int best_child = -1;
  for (int i = 0; i < nd->n_children; ++i) {
    const node_t *ch = &p->a[nd->first_child + i];
    double q = ch->visits == 0 ? 0.0 : ch->value_sum / (double)ch->visits; /* :30-34 */
    volatile float t = 1.5f * ch->prior;
    t = t * (float)sqrt((double)nd->visits);
    t = t / (float)(1 + ch->visits);
    volatile float score = (float)q + t;
    if (score > best) { /* strict >, first max wins */
      best = score;
      best_child = nd->first_child + i;
    }
  }
  return best_child;
}

/* self_play.py:61-68 — idempotent */
static void expand(pool_t *p, int node, const int16_t *moves, int n, const float *priors) {
  if (p->a[node].n_children > 0) return;
  int first = p->n;
  for (int i = 0; i < n; ++i) pool_new(p, node, moves[i], priors[i]);
  p->a[node].first_child = first;
  p->a[node].n_children = n;
}

/* self_play.py:70-80 */
static void update(pool_t *p, int node, double value) {
  while (node >= 0) {
    p->a[node].visits += 1;
    p->a[node].value_sum += value;
    value = -value;
    node = p->a[node].parent;
  }
}

/* self_play.py:156-175 */
static void copy_env(const xqo_state *env, xqo_state *out) {
  xqo_reset(out);
  memcpy(out->board, env->board, XQO_NSQ);
  out->player = env->player;
  out->move_count = env->move_count;
  out->winner = env->winner;
  out->red_king = env->red_king;
  out->black_king = env->black_king;
  out->no_capture = env->no_capture;
}

/* self_play.py:89-154 */
int xqo_mcts_search(const xqo_state *env, int num_simulations, xqo_eval_fn eval,
                    void *ctx, int16_t *root_moves, int32_t *root_visits,
                    int64_t *stats) {
  pool_t pool = {0};
  int root = pool_new(&pool, -1, -1, 0.0f);
  xqo_state *se = (xqo_state *)malloc(sizeof(xqo_state));
  enum { WAVE = 8 }; /* :101 */
  int leaf_node[WAVE], leaf_player[WAVE], leaf_n[WAVE];
  int8_t *leaf_board = (int8_t *)malloc(WAVE * XQO_NSQ);
  int16_t *leaf_moves = (int16_t *)malloc(WAVE * XQO_MAX_MOVES * sizeof(int16_t));
  float *priors = (float *)malloc(WAVE * XQO_MAX_MOVES * sizeof(float));
  double values[WAVE];
  int64_t st[4] = {0, 0, 0, 0};

  for (int start = 0; start < num_simulations; start += WAVE) {
    int count = num_simulations - start < WAVE ? num_simulations - start : WAVE;
    int nq = 0;
    for (int k = 0; k < count; ++k) {
      int node = root;
      copy_env(env, se); /* :115 */
      while (pool.a[node].n_children > 0) { /* :117-119 */
        node = select_child(&pool, node);
        xqo_step_result r;
        xqo_make_move(se, pool.a[node].move, &r);
      }
      int16_t *mv = leaf_moves + nq * XQO_MAX_MOVES;
      int n = xqo_legal_moves(se, mv); /* :123 */
      st[0]++;
      if (n == 0 || se->winner != XQO_WINNER_NONE) { /* :126-135 */
        double v = 0;
        if (se->winner == se->player)
          v = 1;
        else if (se->winner == -se->player)
          v = -1;
        update(&pool, node, v);
        st[3]++;
      } else { /* :138-139 */
        leaf_node[nq] = node;
        leaf_player[nq] = se->player;
        leaf_n[nq] = n;
        memcpy(leaf_board + nq * XQO_NSQ, se->board, XQO_NSQ);
        ++nq;
        st[1]++;
      }
    }
    if (nq > 0) { /* :142-148 */
      eval(ctx, nq, leaf_board, leaf_player, leaf_moves, leaf_n, priors, values);
      st[2]++;
      for (int q = 0; q < nq; ++q) {
        expand(&pool, leaf_node[q], leaf_moves + q * XQO_MAX_MOVES, leaf_n[q],
               priors + q * XQO_MAX_MOVES);
        update(&pool, leaf_node[q], values[q]);
      }
    }
  }
  int nroot = pool.a[root].n_children; /* :151-154 */
  for (int i = 0; i < nroot; ++i) {
    root_moves[i] = (int16_t)pool.a[pool.a[root].first_child + i].move;
    root_visits[i] = pool.a[pool.a[root].first_child + i].visits;
  }
  if (stats) memcpy(stats, st, sizeof(st));
  free(pool.a);
  free(se);
  free(leaf_board);
  free(leaf_moves);
  free(priors);
  return nroot;
}

/* Deterministic evaluator for parity tests and the MCTS CPU baseline.
 * ctx == NULL or *(int*)ctx == 0: hashed priors; *(int*)ctx == 1: flat priors
 * (exercises the first-max tie-break).  Order-independent arithmetic so the
 * CUDA mirror can compute it lane-parallel: integer weights < 2^24, summed
 * exactly in double, one double division, one rounding to float32. */
void xqo_hash_eval(void *ctx, int n_leaves, const int8_t *boards, const int32_t *players,
                   const int16_t *moves, const int32_t *n_moves, float *priors,
                   double *values) {
  int flat = ctx ? *(const int *)ctx : 0;
  for (int q = 0; q < n_leaves; ++q) {
    uint64_t h = xqo_position_hash(boards + q * XQO_NSQ, players[q]);
    const int16_t *mv = moves + q * XQO_MAX_MOVES;
    float *pr = priors + q * XQO_MAX_MOVES;
    int n = n_moves[q];
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
      uint64_t u = flat ? 0 : mix64(h ^ ((uint64_t)(uint16_t)mv[i] * 0x9E3779B97F4A7C15ULL));
      double w = (double)((u >> 40) + 1);
      sum += w;
    }
    for (int i = 0; i < n; ++i) {
      uint64_t u = flat ? 0 : mix64(h ^ ((uint64_t)(uint16_t)mv[i] * 0x9E3779B97F4A7C15ULL));
      double w = (double)((u >> 40) + 1);
      pr[i] = (float)(w / sum);
    }
    values[q] = (double)((h >> 11) & 0xFFFFFu) / 524288.0 - 1.0;
  }
}
