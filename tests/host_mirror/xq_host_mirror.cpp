// tests/host_mirror/xq_host_mirror.cpp — TEST HARNESS ONLY (never shipped, never a fallback).
// Compiles the per-lane device logic of chinesechessai_b200/csrc/xq_rules.cuh (attacked,
// suicide, gen_item, exotic_piece — all XQ_HD) with g++ and runs the warp's work items
// sequentially, so the kernel's legality logic can be fuzzed against the oracle on the CPU.
#include <cstring>

#include "../../chinesechessai_b200/csrc/xq_rules.cuh"

using namespace xq;

static void stage(WarpSmem& w, const int8_t* board) {
  std::memset(&w, 0, sizeof(w));
  std::memcpy(w.sq, board, XQ_NSQ);
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 9; ++c)
      if (w.sq[r * 9 + c] != 0) {
        w.rows[r] |= (uint16_t)(1u << c);
        w.cols[c] |= (uint16_t)(1u << r);
      }
}

alignas(16) static const uint32_t h_leap[kLeapEntries] = {
#include "../../chinesechessai_b200/csrc/xq_leap_table.inc"
};
alignas(16) static const uint16_t h_ray[kRayEntries] = {
#include "../../chinesechessai_b200/csrc/xq_ray_table.inc"
};
alignas(16) static const uint32_t h_knight[kKnightEntries] = {
#include "../../chinesechessai_b200/csrc/xq_knight_table.inc"
};
static const Tables h_tables{h_leap, h_ray, h_knight};
static const uint32_t h_touch[kTouchEntries] = {
#include "../../chinesechessai_b200/csrc/xq_touch_table.inc"
};

// mode 0: every candidate through suicide(); mode 1: the kernel's Phase B (relevance filter,
// sentinel "no move" evaluation, flag bits) executed sequentially.
extern "C" int xqh_legal_moves(const int8_t* board, int player, int red_king, int black_king,
                               int16_t* moves, int* ncand_out, int mode) {
  WarpSmem w;
  stage(w, board);
  Game g{};
  g.player = player;
  g.red_king = red_king;
  g.black_king = black_king;
  const int ownK = player == 1 ? red_king : black_king;
  int n_own = 0, n_kings = 0;
  bool exotic = false;
  for (int s = 0; s < XQ_NSQ; ++s) {
    if ((int)w.sq[s] * player > 0) w.own[n_own++] = (uint8_t)s;
    n_kings += w.sq[s] == player * KING;
    exotic |= exotic_piece(w.sq[s], s, player, ownK < 0 ? 0 : ownK);
  }
  exotic = exotic || !regular_king(w, player, ownK, n_kings);
  int ncand = 0;
  for (int t = 0; t < n_own * 4; ++t) {
    const Item it = gen_item(w, h_tables, player, w.own[t >> 2], t & 3);
    const int cnt = it.empties + (it.e1 >= 0) + (it.e2 >= 0);
    if (ncand + cnt > XQ_CAND_CAP) return -1;
    for (int k = 1; k <= it.empties; ++k) w.cand[ncand++] = (uint16_t)((it.from << 8) | (it.from + k * it.delta));
    if (it.e1 >= 0) w.cand[ncand++] = (uint16_t)((it.from << 8) | it.e1);
    if (it.e2 >= 0) w.cand[ncand++] = (uint16_t)((it.from << 8) | it.e2);
  }
  if (ncand_out) *ncand_out = ncand;
  int n = 0;
  if (mode == 0) {
    for (int j = 0; j < ncand; ++j)
      if (!suicide(w, g, w.cand[j] >> 8, w.cand[j] & 0x7f, true)) {
        if (n < XQ_MAX_MOVES) moves[n] = (int16_t)((w.cand[j] >> 8) * 90 + (w.cand[j] & 0x7f));
        ++n;
      }
    return n;
  }
  const int kr = ownK >= 0 ? ownK / 9 : 0, kc = ownK >= 0 ? ownK - kr * 9 : 0;
  if (mode == 2 && !exotic) {  // the kernel's regular-position path: bitmask test + king probes
    const FastCtx f = make_fast_ctx(w, g, h_touch);
    const bool cur_bad = suicide_fast(f, -1, -1);
    if (ncand_out) ncand_out[1] = 0, ncand_out[2] = 0;
    for (int j = 0; j < ncand; ++j) {
      const int from = w.cand[j] >> 8, to = w.cand[j] & 0x7f;
      bool bad;
      if (from == ownK) {  // king move: 8 probes at the new square + facing, as the king round does
        const Probe p = make_probe(w, to, -player, player, from, to, w.sq[from]);
        bad = false;
        for (int d = 0; d < 4; ++d) bad |= probe_ray(w, p, d);
        for (int d = 0; d < 4; ++d) bad |= probe_diag(w, p, d, false);
        bad |= king_move_facing(g, p, to);
      } else if (touches(kr, kc, from) || touches(kr, kc, to)) {
        bad = suicide_fast(f, from, to);
      } else {
        bad = cur_bad;
      }
      if (!bad) {
        if (n < XQ_MAX_MOVES) moves[n] = (int16_t)(from * 90 + to);
        ++n;
      }
    }
    return n;
  }
  int nwl = 0;
  bool any_irrelevant = false;
  for (int j = 0; j < ncand; ++j) {
    const int c = w.cand[j], from = c >> 8, to = c & 0x7f;
    const bool rel = exotic || from == ownK || touches(kr, kc, from) || touches(kr, kc, to);
    if (rel) w.wl[nwl++] = (uint16_t)j;
    else {
      w.cand[j] = (uint16_t)(c | kCandIrrelevant);
      any_irrelevant = true;
    }
  }
  if (any_irrelevant) w.wl[nwl++] = kWlSentinel;
  if (ncand_out) ncand_out[1] = nwl, ncand_out[2] = exotic;
  bool cur_bad = false;
  for (int i = 0; i < nwl; ++i) {
    const int item = w.wl[i];
    if (item == kWlSentinel) {
      cur_bad = suicide(w, g, -1, -1, exotic);
    } else {
      const int c = w.cand[item];
      if (suicide(w, g, c >> 8, c & 0x7f, exotic)) w.cand[item] = (uint16_t)(c | kCandIllegal);
    }
  }
  for (int j = 0; j < ncand; ++j) {
    const int c = w.cand[j];
    const bool ok = (c & kCandIrrelevant) ? !cur_bad : !(c & kCandIllegal);
    if (ok) {
      if (n < XQ_MAX_MOVES) moves[n] = (int16_t)(((c >> 8) & 0x7f) * 90 + (c & 0x7f));
      ++n;
    }
  }
  return n;
}

// OR of the 8 per-direction probes (what in_check_warp computes with 8 lanes)
extern "C" int xqh_in_check_dirs(const int8_t* board, int player, int current_player, int red_king,
                                 int black_king) {
  WarpSmem w;
  stage(w, board);
  const int K = player == 1 ? red_king : black_king;
  if (K < 0) return 0;
  bool hit = false;
  for (int d = 0; d < 8; ++d) hit |= attacked_dir(w, K, -player, current_player, d);
  return hit ? 1 : 0;
}

// _is_in_check(player) with geometry of `current_player` (chess_env.py:506-548)
extern "C" int xqh_in_check(const int8_t* board, int player, int current_player, int red_king,
                            int black_king) {
  WarpSmem w;
  stage(w, board);
  Game g{};
  g.player = current_player;
  g.red_king = red_king;
  g.black_king = black_king;
  return in_check(w, g, player) ? 1 : 0;
}

extern "C" double xqh_position_change(int type, int player, int from, int to, int enemy_king) {
  return position_change(type, player, from, to, enemy_king);
}

// check_fast(): the check test of make_move for regular positions.  -1 when the position is not
// regular (the kernels then use the general probes).
extern "C" int xqh_check_fast(const int8_t* board, int player, int red_king, int black_king) {
  WarpSmem w;
  stage(w, board);
  Game g{};
  g.player = player;
  g.red_king = red_king;
  g.black_king = black_king;
  const int ownK = player == 1 ? red_king : black_king;
  int n_kings = 0;
  bool exotic = false;
  for (int s = 0; s < XQ_NSQ; ++s) {
    n_kings += w.sq[s] == player * KING;
    exotic |= exotic_piece(w.sq[s], s, player, ownK < 0 ? 0 : ownK);
  }
  if (exotic || !regular_king(w, player, ownK, n_kings)) return -1;
  return check_fast(make_fast_ctx(w, g, h_touch), player) ? 1 : 0;
}

// gen_piece() + gen_dir() (the lane-pair engine's per-piece generator) against gen_item() on every
// square that holds a piece of `player` (any code, poked boards included).  Returns the number of
// (square, slot) items that differ.
extern "C" int xqh_gen_piece_mismatches(const int8_t* board, int player) {
  WarpSmem w;
  stage(w, board);
  int bad = 0;
  for (int s = 0; s < XQ_NSQ; ++s) {
    if ((int)w.sq[s] * player <= 0) continue;
    const PieceGen pg = gen_piece(w, h_tables, player, s);
    for (int d = 0; d < 4; ++d) {
      const Item a = gen_item(w, h_tables, player, s, d), b = gen_dir(w, pg, player, d);
      bad += a.from != b.from || a.empties != b.empties || a.e1 != b.e1 || a.e2 != b.e2 ||
             (a.empties > 0 && a.delta != b.delta);
    }
  }
  return bad;
}

// king_move_fast() against the general suicide() for every candidate of the own king on a regular
// position.  Returns the number of candidates that differ, -1 when the position is not regular.
extern "C" int xqh_king_move_mismatches(const int8_t* board, int player, int red_king, int black_king) {
  WarpSmem w;
  stage(w, board);
  Game g{};
  g.player = player;
  g.red_king = red_king;
  g.black_king = black_king;
  const int ownK = player == 1 ? red_king : black_king, ek = player == 1 ? black_king : red_king;
  int n_kings = 0;
  bool exotic = false;
  for (int s = 0; s < XQ_NSQ; ++s) {
    n_kings += w.sq[s] == player * KING;
    exotic |= exotic_piece(w.sq[s], s, player, ownK < 0 ? 0 : ownK);
  }
  if (exotic || !regular_king(w, player, ownK, n_kings)) return -1;
  int bad = 0;
  for (int d = 0; d < 4; ++d) {
    const Item it = gen_item(w, h_tables, player, ownK, d);
    const int ts[2] = {it.e1, it.e2};
    for (int k = 0; k < 2; ++k)
      if (ts[k] >= 0)
        bad += king_move_fast(w, h_tables, player, ownK, ts[k], ek) != suicide(w, g, ownK, ts[k], false);
  }
  return bad;
}
