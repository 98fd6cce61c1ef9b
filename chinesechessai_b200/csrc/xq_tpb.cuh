// xq_tpb.cuh — thread-per-board ("tile width 1") rules engine for the fused playout.
//
// The warp-per-board engine (xq_rules.cuh) keeps ~20 of 32 lanes busy and spends much of its
// issue budget on warp-collective glue.  For the fused random playout, where 65,536 boards are
// in flight and nothing has to be cooperative, one THREAD per board is the better mapping: the
// same per-lane functions (gen_item, suicide_fast, make_fast_ctx, attacked, position_change —
// all shared with the warp engine and the host mirror) run sequentially per board, every issued
// instruction serves up to 32 boards, and the only inefficiency left is divergence between the
// boards of a warp (different piece types / list lengths at the same loop index).
// Semantics and outputs are identical to the warp engine (same tests, same goldens).
#pragma once

#include "xq_rules.cuh"

#if defined(__CUDACC__)
namespace xq {

constexpr int kTpbOwnCap = 32;
constexpr int kTpbMoveCap = XQ_MAX_MOVES;  // candidates and legal moves share this array

// Per-thread slab in shared memory.  107 words: an odd word stride keeps warp-uniform offsets
// free of bank conflicts.
struct ThreadBoard {
  int8_t sq[XQ_BOARD_STRIDE];
  uint16_t rows[10];
  uint16_t cols[10];
  uint8_t own[kTpbOwnCap];
  uint16_t mv[kTpbMoveCap];  // from<<8 | to
  uint8_t n_kings[2];  // red / black king pieces on the board (constant while a game runs:
                       // capturing a king ends it, chess_env.py:291-299)
  uint16_t pad;
};
static_assert(sizeof(ThreadBoard) == 428 && (sizeof(ThreadBoard) / 4) % 2 == 1, "ThreadBoard stride");

// Two adjacent lanes that share one board (xq_pair.cuh).  Collectives use the pair's own mask,
// so the pairs of a warp may diverge freely.
struct Pair {
  static __device__ __forceinline__ int sub() { return (int)(threadIdx.x & 1u); }
  static __device__ __forceinline__ unsigned mask() { return 3u << (threadIdx.x & 30u); }
  static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
  template <class T>
  static __device__ __forceinline__ T other(T v) { return __shfl_xor_sync(mask(), v, 1); }
};

__device__ __forceinline__ int tpb_packed(unsigned c) { return (int)(c >> 8) * 90 + (int)(c & 0x7fu); }

__device__ __forceinline__ void tpb_load(ThreadBoard& w, const int8_t* __restrict__ row) {
#pragma unroll
  for (int i = 0; i < XQ_BOARD_STRIDE / 4; ++i)
    reinterpret_cast<uint32_t*>(w.sq)[i] = reinterpret_cast<const uint32_t*>(row)[i];
#pragma unroll 1
  for (int r = 0; r < 10; ++r) {
    unsigned m = 0;
    for (int c = 0; c < 9; ++c) m |= (w.sq[r * 9 + c] != 0 ? 1u : 0u) << c;
    w.rows[r] = (uint16_t)m;
  }
#pragma unroll 1
  for (int c = 0; c < 9; ++c) {
    unsigned m = 0;
    for (int r = 0; r < 10; ++r) m |= (w.sq[r * 9 + c] != 0 ? 1u : 0u) << r;
    w.cols[c] = (uint16_t)m;
  }
  int nr = 0, nb = 0;
#pragma unroll 1
  for (int s = 0; s < XQ_NSQ; ++s) {
    nr += w.sq[s] == KING;
    nb += w.sq[s] == -KING;
  }
  w.n_kings[0] = (uint8_t)(nr > 255 ? 255 : nr);
  w.n_kings[1] = (uint8_t)(nb > 255 ? 255 : nb);
}

__device__ __forceinline__ uint64_t tpb_board_key(const ThreadBoard& w) {
  uint64_t h = 0;
#pragma unroll 1
  for (int s = 0; s < XQ_NSQ; ++s) {
    const int p = w.sq[s];
    if (p != 0) h ^= piece_key(p, s);
  }
  return h;
}

// get_legal_moves (chess_env.py:76-121), sequential.  Legal moves end up in w.mv[0..n) in the
// reference's order.  *checked (optional) = make_move's is_checking for the move that led here
// (:317: this side's king attacked under the previous mover's geometry).
__device__ __forceinline__ int tpb_movegen(ThreadBoard& w, Game& g, const uint32_t* __restrict__ leap,
                                           bool* checked) {
  const int player = g.player;
  const int ownK = player == 1 ? g.red_king : g.black_king;
  // :82-87 scan order, four squares per word: sign/zero tests on packed bytes give the own
  // pieces (any code of the mover's sign) and the enemy K/A/B (|code| <= 3) of exotic_piece()
  int n_own = 0;
  bool ex = false;
  int lo, hi;
  exotic_window(player, ownK < 0 ? 0 : ownK, &lo, &hi);
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(w.sq);
#pragma unroll 1
  for (int i = 0; i < 23; ++i) {
    uint32_t x = sw[i];
    if (i == 22) x &= 0xFFFFu;  // squares 88, 89; the row padding is not part of the board
    const uint32_t l7 = x & 0x7F7F7F7Fu;
    const uint32_t neg = x & 0x80808080u;
    const uint32_t pos = (l7 + 0x7F7F7F7Fu) & ~x & 0x80808080u;
    uint32_t own = player == 1 ? pos : neg;
    const uint32_t kab = player == 1 ? ((l7 + 0x03030303u) & neg) : (pos & ~(l7 + 0x7C7C7C7Cu));
    if (kab) {  // a word spans 4 squares, the window >= 27: testing both ends is exact
      const int a = 4 * i + ((__ffs(kab) - 1) >> 3), b = 4 * i + ((31 - __clz(kab)) >> 3);
      ex |= (a >= lo && a < hi) || (b >= lo && b < hi);
    }
    while (own) {
      const int s = 4 * i + ((__ffs(own) - 1) >> 3);
      own &= own - 1;
      if (n_own < kTpbOwnCap) w.own[n_own] = (uint8_t)s;
      ++n_own;
    }
  }
  if (n_own > kTpbOwnCap) {
    n_own = kTpbOwnCap;
    g.flags |= XQ_F_OVERFLOW;
  }
  const bool exotic = ex || !regular_king(w, player, ownK, (int)w.n_kings[player == 1 ? 0 : 1]);

  // candidates in generator order
  int nc = 0;
  const Tables tb{leap, g_ray};
#pragma unroll 1
  for (int t = 0; t < n_own * 4; ++t) {
    const Item it = gen_item(w, tb, player, w.own[t >> 2], t & 3);
    const int cnt = it.empties + (it.e1 >= 0) + (it.e2 >= 0);
    if (nc + cnt > kTpbMoveCap) {
      g.flags |= XQ_F_OVERFLOW;
      break;
    }
    const unsigned fs = (unsigned)it.from << 8;
    uint16_t* out = &w.mv[nc];
    unsigned v = fs | (unsigned)it.from;  // from + k*delta stays inside the low byte
#pragma unroll 1
    for (int k = it.empties; k > 0; --k) *out++ = (uint16_t)(v += (unsigned)it.delta);
    if (it.e1 >= 0) *out++ = (uint16_t)(fs | (unsigned)it.e1);
    if (it.e2 >= 0) *out = (uint16_t)(fs | (unsigned)it.e2);
    nc += cnt;
  }

  // legality (:118), compacting in place
  int n = 0;
  if (!exotic) {
    const FastCtx f = make_fast_ctx(w, g, g_touch);
    int kfirst = -1, kcount = 0;
#pragma unroll 1
    for (int j = 0; j < nc; ++j) {  // non-king moves: bitmask test, uniform code for all lanes
      const unsigned c = w.mv[j];
      const int from = (int)(c >> 8), to = (int)(c & 0x7fu);
      if (from == ownK) {  // keep the king's candidates (contiguous) in place for the probe loop
        if (kcount == 0) kfirst = n;
        ++kcount;
        w.mv[n++] = (uint16_t)(c | kCandIllegal);  // provisional
      } else if (!suicide_fast(f, from, to)) {
        w.mv[n++] = (uint16_t)c;
      }
    }
    // king moves: general probes at the new square + kings facing (:448-451)
    int removed = 0;
#pragma unroll 1
    for (int k = 0; k < kcount; ++k) {
      const unsigned c = w.mv[kfirst + k] & 0x7fffu;
      const bool bad = suicide(w, g, (int)(c >> 8), (int)(c & 0x7fu), false);
      if (!bad) w.mv[kfirst + k - removed] = (uint16_t)c;
      else ++removed;
    }
    if (removed) {  // close the gap behind the king's block
#pragma unroll 1
      for (int j = kfirst + kcount; j < n; ++j) w.mv[j - removed] = w.mv[j];
      n -= removed;
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < nc; ++j) {
      const unsigned c = w.mv[j];
      if (!suicide(w, g, (int)(c >> 8), (int)(c & 0x7fu), true)) w.mv[n++] = (uint16_t)c;
    }
  }
  if (checked) *checked = ownK >= 0 && attacked(w, ownK, -player, -player, -1, -1, 0, true, nullptr);
  return n;
}

struct TpbStep {
  double reward;
  int is_int, done, from, to, moving, captured;
  uint64_t key_next;
};

// make_move part 1 (chess_env.py:253-314,:338,:348-349), sequential twin of step_apply<L>.
// PAIR: both lanes of a pair keep the game registers; lane 0 alone writes the shared slab and
// the history.
template <bool PAIR = false>
__device__ __forceinline__ TpbStep tpb_apply(ThreadBoard& w, Game& g, int from, int to,
                                             uint64_t* __restrict__ hist, int hist_cap) {
  TpbStep o;
  const int captured = w.sq[to], moving = w.sq[from];
  if (PAIR) Pair::sync();  // both lanes have read the old squares
  if (!PAIR || Pair::sub() == 0) {
    w.sq[to] = (int8_t)moving;
    w.sq[from] = 0;
    const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
    w.rows[fr] &= ~(1u << fc);
    w.cols[fc] &= ~(1u << fr);
    if (moving != 0) {
      w.rows[tr] |= (uint16_t)(1u << tc);
      w.cols[tc] |= (uint16_t)(1u << tr);
    } else {
      w.rows[tr] &= ~(1u << tc);
      w.cols[tc] &= ~(1u << tr);
    }
  }
  if (PAIR) Pair::sync();
  if (moving != 0) g.bkey ^= piece_key(moving, from) ^ piece_key(moving, to);
  if (captured != 0) g.bkey ^= piece_key(captured, to);
  if (moving == KING) g.red_king = to;
  else if (moving == -KING) g.black_king = to;
  if (captured == KING) g.red_king = -1;
  else if (captured == -KING) g.black_king = -1;
  g.no_capture = captured != 0 ? 0 : g.no_capture + 1;
  o.from = from; o.to = to; o.moving = moving; o.captured = captured;
  o.reward = 0.0;
  o.is_int = 1;
  o.done = 0;
  const int acap = captured < 0 ? -captured : captured;
  if (acap == KING) {
    g.winner = g.player;
    o.reward = 100.0;
    o.done = 1;
    g.reason = XQ_REASON_KING_CAPTURE;
    g.done = 1;
  } else if (captured != 0) {
    const double base = acap == ROOK ? 9.0 : acap == CANNON ? 4.5 : acap == KNIGHT ? 4.0
                        : (acap == BISHOP || acap == ADVISOR) ? 2.0 : acap == PAWN ? 1.0 : 0.0;
    o.reward = xq_dmul(base, 2.0);
    o.is_int = 0;
    if (acap == ADVISOR || acap == BISHOP) o.reward = xq_dadd(o.reward, 3.0);
  }
  if (g.hist_len < hist_cap) {
    if (!PAIR || Pair::sub() == 0) hist[g.hist_len] = g.bkey ^ side_key(g.player);
    g.hist_len += 1;
  } else {
    g.flags |= XQ_F_OVERFLOW;
  }
  g.player = -g.player;
  g.move_count += 1;
  o.key_next = g.bkey ^ side_key(g.player);
  return o;
}

// make_move part 2 (:318-345, :352-404), sequential twin of step_finish<L>.
// HIST_CG: read the history through L2 (the queue-fed kernel, where earlier entries were appended
// by warps on other SMs); otherwise plain loads (entries written by this thread or, in the
// SM-scheduled kernel, by other warps of the same CTA behind a CTA-scope fence).
template <bool PAIR = false, bool HIST_CG = false>
__device__ __forceinline__ void tpb_finish(const ThreadBoard& w, Game& g, TpbStep& o, int n_legal,
                                           bool checking, const uint64_t* __restrict__ hist) {
  const int mover = -g.player;
  if (!o.done && checking) {
    if (g.cchecks == 0) { o.reward = xq_dadd(o.reward, 15.0); o.is_int = 0; }
    else if (g.cchecks == 1) { o.reward = xq_dadd(o.reward, 10.0); o.is_int = 0; }
    else if (g.cchecks == 2) { o.reward = xq_dadd(o.reward, 5.0); o.is_int = 0; }
    g.cchecks += 1;
  } else {
    g.cchecks = 0;
    if (o.captured == 0 && !o.done) {
      const int ek = mover == 1 ? g.black_king : g.red_king;
      const double pcg = position_change(o.moving < 0 ? -o.moving : o.moving, mover, o.from, o.to, ek);
      o.reward = xq_dadd(o.reward, xq_dmul(pcg, 0.01));
      o.is_int = 0;
    }
  }
  g.check_bits = (g.check_bits << 1) | (checking ? 1u : 0u);
  g.check_len += 1;
  if (o.done) return;
  const bool chk_now = n_legal == 0 ? in_check(w, g, g.player) : false;
  if (n_legal == 0 && chk_now) {
    o.done = 1; o.reward = 200.0; o.is_int = 1;
    g.winner = -g.player;
    g.reason = XQ_REASON_CHECKMATE;
  } else {
    int cnt = 0;
#pragma unroll 1
    for (int i = PAIR ? Pair::sub() : 0; i < g.hist_len; i += PAIR ? 2 : 1)
      cnt += (HIST_CG ? __ldcg(hist + i) : hist[i]) == o.key_next;
    if (PAIR) cnt += Pair::other(cnt);
    if (cnt >= 3) {
      o.done = 1; o.reward = 0.0; o.is_int = 1;
      g.winner = 0;
      g.reason = XQ_REASON_REPETITION;
    } else if (g.no_capture >= 100) {
      o.done = 1; o.reward = 0.0; o.is_int = 1;
      g.winner = 0;
      g.reason = XQ_REASON_FIFTY;
    } else if (n_legal == 0) {
      o.done = 1; o.reward = 100.0; o.is_int = 1;
      g.winner = -g.player;
      g.reason = XQ_REASON_STALEMATE;
    } else if (g.check_len >= 12 && __popc(g.check_bits & 0xFFFu) >= 10) {
      o.done = 1; o.reward = -10.0; o.is_int = 1;
      g.winner = -g.player;
      g.reason = XQ_REASON_PERPETUAL_CHECK;
    }
  }
  if (!o.done && g.move_count >= 70) {
    o.done = 1; o.reward = -2.0; o.is_int = 1;
    g.winner = 0;
    g.reason = XQ_REASON_MOVE_CAP;
  }
  if (o.done) g.done = 1;
}

// shared pick rule: index into w.mv[0..n)
__device__ __forceinline__ int tpb_pick(const ThreadBoard& w, int n, uint64_t seed, uint32_t game_id,
                                        uint32_t ply, int capture_bias) {
  uint32_t x[4];
  philox4x32(game_id, ply, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  if (capture_bias > 0 && (int)(x[1] & 0xFFu) < capture_bias) {
    int ncap = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) ncap += w.sq[w.mv[i] & 0x7fu] != 0;
    if (ncap > 0) {
      int k = (int)(x[0] % (uint32_t)ncap);
#pragma unroll 1
      for (int i = 0; i < n; ++i)
        if (w.sq[w.mv[i] & 0x7fu] != 0 && k-- == 0) return i;
    }
  }
  return (int)(x[0] % (uint32_t)n);
}

}  // namespace xq
#endif  // __CUDACC__
