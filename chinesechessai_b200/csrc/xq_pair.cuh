// xq_pair.cuh — two lanes per board: fused playout, xq_step and xq_legal_moves on large batches.
//
// The thread-per-board engine (xq_tpb.cuh) issues the fewest instructions per ply, but at the
// cfg 2 size (65,536 boards = 2,048 warps = 3.5 warps per scheduler) it is bound by dependent-
// issue latency: issue-active 52 % (profiles/).  Here two ADJACENT lanes share one board slab and
// split the divisible work of a ply, so the same batch runs as twice as many warps with about
// half the dependent chain per lane:
//   * own-piece scan: lane 0 the squares of words 0..11, lane 1 words 12..22 (four squares per
//     32-bit load, branch-free append), each into its half of the own list;
//   * candidate generation: by halves of the piece list, one PIECE per iteration (gen_piece /
//     gen_dir: type, row / column and the four table words derived once, four slots unrolled);
//     lane 0's candidates grow up from mv[0], lane 1's down from mv[127]; a quiet ray is one
//     descriptor expanded by a flat pass;
//   * legality: suicide_fast over EQUAL halves of the whole candidate list; make_move's check
//     test from the same masks (check_fast); the king's own moves by table (king_move_fast), the
//     lanes taking alternating candidates; the general probes only on irregular (poked) boards;
//   * the fused playouts (pair_movegen<LAZY>) keep the verdicts in a per-lane bit mask, sum the
//     digest's list term inside the legality loop and pick by rank in the mask — no compaction;
//     the API kernels and the trace mode mark in place and compact each lane's run;
//   * the repetition scan by alternating history entries.
// Bookkeeping (make_move's scalar part, the pick) is replicated in both lanes' registers; lane 0
// writes the slab.  The legal list in the reference's order is lane 0's run followed by lane
// 1's and is never merged physically (pair_move_at).  Pair-mask shuffles and __syncwarp only.
// Outputs are bit-identical to the other two engines (same tests, same digests).
#pragma once

#include "xq_tpb.cuh"

#if defined(__CUDACC__)
namespace xq {

// i-th legal move of the pair's list (n0 = lane 0's count)
__device__ __forceinline__ unsigned pair_move_at(const ThreadBoard& w, int i, int n0) {
  return w.mv[i < n0 ? i : kTpbMoveCap - 1 - (i - n0)];
}

// The legal list WITHOUT compaction (fused playouts): the candidates stay where move generation
// put them, each lane keeps the verdicts of its half of the list (entries j0 .. of the pair's
// numbering: lane 0's run, then lane 1's) as a bit mask.  The digest term is summed inside the
// legality loop and the one move a ply needs is found by rank in the mask, so the per-slot
// compaction pass (8 % of the instructions of a ply, at 17 of 32 lanes) is not run at all.
struct PairLegal {
  unsigned long long mask;  // bit t: entry j0 + t is legal
  int j0;                   // first entry of this lane's half
  int nc0;                  // length of lane 0's run (address of entry j: pair_entry_addr)
  int n_a;                  // legal entries in lane 0's half
};
__device__ __forceinline__ int pair_entry_addr(int j, int nc0) {
  return j < nc0 ? j : kTpbMoveCap - 1 - (j - nc0);
}
// position of the k-th (0-based) set bit
__device__ __forceinline__ int nth_bit64(unsigned long long m, int k) {
  const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
  const int cl = __popc(lo);
  return k < cl ? (int)__fns(lo, 0u, k + 1) : 32 + (int)__fns(hi, 0u, k - cl + 1);
}

__device__ __forceinline__ void pair_load(ThreadBoard& w, const int8_t* __restrict__ row) {
  const int sub = Pair::sub();
#pragma unroll
  for (int i = 0; i < XQ_BOARD_STRIDE / 8; ++i)
    reinterpret_cast<uint32_t*>(w.sq)[2 * i + sub] = reinterpret_cast<const uint32_t*>(row)[2 * i + sub];
  Pair::sync();
  if (sub == 0) {
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
      unsigned m = 0;
      for (int c = 0; c < 9; ++c) m |= (w.sq[r * 9 + c] != 0 ? 1u : 0u) << c;
      w.rows[r] = (uint16_t)m;
    }
  } else {
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
  Pair::sync();
}

__device__ __forceinline__ uint64_t pair_board_key(const ThreadBoard& w) {
  uint64_t h = 0;
#pragma unroll 1
  for (int s = Pair::sub(); s < XQ_NSQ; s += 2) {
    const int p = w.sq[s];
    if (p != 0) h ^= piece_key(p, s);
  }
  return h ^ (uint64_t)Pair::other((unsigned long long)h);
}

// get_legal_moves (chess_env.py:76-121) by a pair.  Returns the total count; n_first = lane 0's
// share, checked as in tpb_movegen, lsum = sum over the list of (move_i + 1) * (2 i + 1) mod 2^32
// (the digest's list term, folded into the compaction loop).
template <bool LAZY = false>
__device__ __forceinline__ int pair_movegen(ThreadBoard& w, Game& g, const uint32_t* __restrict__ leap,
                                            bool want_check, bool& checked, int& n_first,
                                            unsigned& lsum, PairLegal* lazy = nullptr) {
  constexpr int kHalfOwn = kTpbOwnCap / 2;
  const int sub = Pair::sub();
  const int player = g.player;
  const int ownK = player == 1 ? g.red_king : g.black_king;
  // Own pieces in scan order (:82-87) and the exotic_piece() hint, four squares per word as in
  // tpb_movegen; lane 0 scans words 0..11 into own[0..), lane 1 words 12..22 into own[16..).
  int my_own = 0;
  int ex = 0;
  {
    int lo, hi;
    exotic_window(player, ownK < 0 ? 0 : ownK, &lo, &hi);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(w.sq);
    const int i_end = sub ? 23 : 12;
#pragma unroll 1
    for (int i = sub ? 12 : 0; i < i_end; ++i) {
      uint32_t x = sw[i];
      if (i == 22) x &= 0xFFFFu;  // squares 88, 89; the row padding is not part of the board
      const uint32_t l7 = x & 0x7F7F7F7Fu;
      const uint32_t neg = x & 0x80808080u;
      const uint32_t pos = (l7 + 0x7F7F7F7Fu) & ~x & 0x80808080u;
      uint32_t own = player == 1 ? pos : neg;
      const uint32_t kab = player == 1 ? ((l7 + 0x03030303u) & neg) : (pos & ~(l7 + 0x7C7C7C7Cu));
      if (kab) {  // a word spans 4 squares, the window >= 27: testing both ends is exact
        const int a = 4 * i + ((__ffs(kab) - 1) >> 3), b = 4 * i + ((31 - __clz(kab)) >> 3);
        ex |= ((a >= lo && a < hi) || (b >= lo && b < hi)) ? 1 : 0;
      }
      // branch-free append: every square is written at the cursor, the cursor moves on only
      // past an own piece (so the lanes of a warp stay converged whatever the piece counts)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (my_own < kHalfOwn) w.own[sub * kHalfOwn + my_own] = (uint8_t)(4 * i + b);
        my_own += (int)((own >> (8 * b + 7)) & 1u);
      }
    }
  }
  if (my_own > kHalfOwn) {
    my_own = kHalfOwn;
    g.flags |= XQ_F_OVERFLOW;
  }
  ex |= Pair::other(ex);
  const int other_own = Pair::other(my_own);
  const int own0 = sub ? other_own : my_own, n_own = my_own + other_own;
  const bool exotic = ex != 0 || !regular_king(w, player, ownK, (int)w.n_kings[player == 1 ? 0 : 1]);
  Pair::sync();

  // candidates: lane 0 takes the first half of the pieces, lane 1 the rest; the own king's
  // block of candidates (contiguous) is noted on the way
  const int dir = sub ? -1 : 1, base = sub ? kTpbMoveCap - 1 : 0;
  const Tables tb{leap, g_ray, g_knight};
  // one piece per iteration, its four generator slots unrolled (gen_piece / gen_dir)
  const int p_half = n_own >> 1, p_end = sub ? n_own : p_half;
  int nc = 0, kfirst = 0, kcount = 0;
#pragma unroll 1
  for (int pi = sub ? p_half : 0; pi < p_end; ++pi) {
    const int from = w.own[pi < own0 ? pi : kHalfOwn + (pi - own0)];
    const PieceGen pg = gen_piece(w, tb, player, from);
    const int nc_piece = nc;
    // a piece has at most 17 candidates (a rook with both lines open); only a list that close to
    // its capacity needs the exact per-slot accounting
    const bool tight = nc + 17 > kTpbMoveCap;
    const unsigned fs = (unsigned)from << 8;
    uint16_t* out = &w.mv[base + dir * nc];
    bool full = false;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const Item it = gen_dir(w, pg, player, d);
      const int cnt = it.empties + (it.e1 >= 0) + (it.e2 >= 0);
      if (tight && nc + cnt > kTpbMoveCap) {
        g.flags |= XQ_F_OVERFLOW;
        full = true;
        break;
      }
      // a quiet ray reserves its `empties` slots and leaves ONE descriptor in the first of them
      // (from<<8 | 0x80 | d<<5 | empties); the flat pass below expands it in place
      if (it.empties > 0) *out = (uint16_t)(fs | 0x80u | ((unsigned)d << 5) | (unsigned)it.empties);
      out += dir * it.empties;
      if (it.e1 >= 0) { *out = (uint16_t)(fs | (unsigned)it.e1); out += dir; }
      if (it.e2 >= 0) { *out = (uint16_t)(fs | (unsigned)it.e2); out += dir; }
      nc += cnt;
    }
    if (from == ownK && nc > nc_piece) {
      if (kcount == 0) kfirst = nc_piece;
      kcount += nc - nc_piece;
    }
    if (full) break;
  }
  {  // expand the ray descriptors: one slot per iteration, all lanes in step (a branch-free form
     // of this loop — selects instead of the three paths — measured 1 % slower)
    unsigned v = 0, delta = 0;
    int left = 0;
#pragma unroll 1
    for (int j = 0; j < nc; ++j) {
      uint16_t* slot = &w.mv[base + dir * j];
      if (left > 0) {
        *slot = (uint16_t)(v += delta);
        --left;
      } else {
        const unsigned c = *slot;
        if (c & 0x80u) {  // slot d of gen_item: (0,1),(0,-1),(1,0),(-1,0)
          delta = ((c & 0x40u) ? 9u : 1u) * ((c & 0x20u) ? 0xFFFFFFFFu : 1u);
          left = (int)(c & 15u) - 1;
          v = (c & 0x7f00u) | ((c >> 8) + delta);  // from + delta stays inside the low byte
          *slot = (uint16_t)v;
        }
      }
    }
  }
  // the two parts share the array: more than kTpbMoveCap candidates in total is an overflow
  // (every entry is still some valid candidate, so nothing downstream can go out of range)
  int nc_other = Pair::other(nc);
  if (nc + nc_other > kTpbMoveCap) {
    g.flags |= XQ_F_OVERFLOW;
    if (sub == 0) nc = kTpbMoveCap - nc_other;
    else nc_other = kTpbMoveCap - nc;
  }
  const int nc0 = sub ? nc_other : nc, total = nc + nc_other;
  // the candidate list as a whole: entry j of lane 0's run, then of lane 1's
#define XQ_PAIR_ADDR(j) ((j) < nc0 ? (j) : kTpbMoveCap - 1 - ((j) - nc0))
  int k_total = 0, k_start = 0;  // the king's block in that numbering (regular boards)
  if (!exotic) {
    const int kinfo = kcount > 0 ? (kcount << 8) | (sub ? nc0 + kfirst : kfirst) : 0;
    const int kboth = kinfo | Pair::other(kinfo);  // at most one lane holds the king
    k_total = kboth >> 8;
    k_start = kboth & 0xff;
  }
  Pair::sync();

  if constexpr (LAZY) {
    // legality (:118) + the digest's list term in ONE pass over equal halves of the list; the
    // verdicts go into a register mask, the list is left as it is
    int chk = 0, n = 0;
    unsigned s0 = 0, s1 = 0;
    unsigned long long legal = 0, bit = 1;
    const int cut = (total + 1) >> 1, j0 = sub ? cut : 0, j_end = sub ? total : cut;
    if (!exotic) {
      const FastCtx f = make_fast_ctx(w, g, g_touch);
      if (want_check) chk = check_fast(f, player) ? 1 : 0;
      // the king's own moves first (:448-451, table test): marked in place, the lanes take
      // alternating candidates
      const int ek = player == 1 ? g.black_king : g.red_king;
#pragma unroll 1
      for (int q = sub; q < k_total; q += 2) {
        const int a = XQ_PAIR_ADDR(k_start + q);
        const unsigned c = w.mv[a];
        if (king_move_fast(w, tb, player, ownK, (int)(c & 0x7fu), ek)) w.mv[a] = (uint16_t)(c | kCandIllegal);
      }
      Pair::sync();
#pragma unroll 1
      for (int j = j0; j < j_end; ++j, bit <<= 1) {
        const unsigned c = w.mv[XQ_PAIR_ADDR(j)];
        const int from = (int)((c >> 8) & 0x7fu), to = (int)(c & 0x7fu);
        const bool bad = from == ownK ? (c & kCandIllegal) != 0 : suicide_fast(f, from, to);
        if (!bad) {
          legal |= bit;
          const unsigned p1 = (unsigned)tpb_packed(c) + 1u;
          s0 += p1;
          s1 += p1 * (unsigned)(2 * n + 1);
          ++n;
        }
      }
    } else {
      // irregular boards: the general test on the lane's own run (marks), then the same pass
#pragma unroll 1
      for (int j = 0; j < nc; ++j) {
        const unsigned c = w.mv[base + dir * j];
        if (suicide(w, g, (int)(c >> 8), (int)(c & 0x7fu), true)) w.mv[base + dir * j] = (uint16_t)(c | kCandIllegal);
      }
      if (want_check && ownK >= 0 && sub == 0)
        chk = attacked(w, ownK, -player, -player, -1, -1, 0, true, nullptr) ? 1 : 0;
      Pair::sync();
#pragma unroll 1
      for (int j = j0; j < j_end; ++j, bit <<= 1) {
        const unsigned c = w.mv[XQ_PAIR_ADDR(j)];
        if (!(c & kCandIllegal)) {
          legal |= bit;
          const unsigned p1 = (unsigned)tpb_packed(c) + 1u;
          s0 += p1;
          s1 += p1 * (unsigned)(2 * n + 1);
          ++n;
        }
      }
    }
    chk |= Pair::other(chk);
    checked = chk != 0;
    const int n_other = Pair::other(n);
    n_first = sub ? n_other : n;
    const unsigned part = sub ? s1 + 2u * (unsigned)n_first * s0 : s1;  // lane 1's i = n_a + local i
    lsum = part + Pair::other(part);
    g.flags |= Pair::other(g.flags);
    lazy->mask = legal;
    lazy->j0 = j0;
    lazy->nc0 = nc0;
    lazy->n_a = n_first;
    Pair::sync();
    return n + n_other;
  }
  // legality (:118): verdicts are MARKED in place (kCandIllegal), nothing moves yet
  int chk = 0;
  if (!exotic) {
    // non-king moves, bitmask test: the list is cut in two equal runs, so the lanes' trip counts
    // match however the pieces fell (a rook's 17 candidates vs a pawn's 1)
    const FastCtx f = make_fast_ctx(w, g, g_touch);
    const int cut = (total + 1) >> 1, j_end = sub ? total : cut;
#pragma unroll 1
    for (int j = sub ? cut : 0; j < j_end; ++j) {
      const int a = XQ_PAIR_ADDR(j);
      const unsigned c = w.mv[a];
      const int from = (int)(c >> 8), to = (int)(c & 0x7fu);
      if (from != ownK && suicide_fast(f, from, to)) w.mv[a] = (uint16_t)(c | kCandIllegal);
    }
    // make_move's check test from the same masks (both lanes, no divergence); the probe round
    // below then only has the king's own moves left
    if (want_check) chk = check_fast(f, player) ? 1 : 0;
    // the king's own moves (:448-451): table test, the lanes take alternating candidates
    const int ek = player == 1 ? g.black_king : g.red_king;
#pragma unroll 1
    for (int q = sub; q < k_total; q += 2) {
      const int a = XQ_PAIR_ADDR(k_start + q);
      const unsigned c = w.mv[a];
      if (king_move_fast(w, tb, player, ownK, (int)(c & 0x7fu), ek)) w.mv[a] = (uint16_t)(c | kCandIllegal);
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < nc; ++j) {  // irregular boards: the general test, king moves included
      const unsigned c = w.mv[base + dir * j];
      if (suicide(w, g, (int)(c >> 8), (int)(c & 0x7fu), true)) w.mv[base + dir * j] = (uint16_t)(c | kCandIllegal);
    }
  }
  // irregular boards: make_move's check test (:317) by the general probes — the king where it
  // stands, under the previous mover's geometry, with the K/A/B probes on (A.3)
  if (exotic && want_check && ownK >= 0 && sub == 0)
    chk = attacked(w, ownK, -player, -player, -1, -1, 0, true, nullptr) ? 1 : 0;
#undef XQ_PAIR_ADDR
  chk |= Pair::other(chk);  // also orders the marks before the compaction reads
  checked = chk != 0;
  Pair::sync();

  // each lane compacts its own run in place and sums its part of the digest term
  int n = 0;
  unsigned s0 = 0, s1 = 0;
#pragma unroll 1
  for (int j = 0; j < nc; ++j) {
    const unsigned c = w.mv[base + dir * j];
    if (!(c & kCandIllegal)) {
      w.mv[base + dir * n] = (uint16_t)c;
      const unsigned p1 = (unsigned)tpb_packed(c) + 1u;
      s0 += p1;
      s1 += p1 * (unsigned)(2 * n + 1);
      ++n;
    }
  }
  const int n_other = Pair::other(n);
  n_first = sub ? n_other : n;
  const unsigned part = sub ? s1 + 2u * (unsigned)n_first * s0 : s1;  // lane 1's i = n0 + local i
  lsum = part + Pair::other(part);
  g.flags |= Pair::other(g.flags);
  Pair::sync();
  return n + n_other;
}

// shared pick rule on the uncompacted list: returns the picked entry (from << 8 | to)
// in both lanes.  The i-th legal move is the i-th set bit of the two lanes' masks taken together.
__device__ __forceinline__ unsigned pair_pick_lazy(const ThreadBoard& w, const PairLegal& L, int n,
                                                   uint64_t seed, uint32_t game_id, uint32_t ply,
                                                   int capture_bias) {
  const int sub = Pair::sub();
  uint32_t x[4];
  philox4x32(game_id, ply, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  unsigned mine = 0;
  int owner = -1;
  if (capture_bias > 0 && (int)(x[1] & 0xFFu) < capture_bias) {
    int cap = 0;
#pragma unroll 1
    for (unsigned long long m = L.mask; m; m &= m - 1) {
      const unsigned c = w.mv[pair_entry_addr(L.j0 + __ffsll((long long)m) - 1, L.nc0)];
      cap += w.sq[c & 0x7fu] != 0;
    }
    const int cap_o = Pair::other(cap), cap_a = sub ? cap_o : cap, ncap = cap + cap_o;
    if (ncap > 0) {
      const int k = (int)(x[0] % (uint32_t)ncap);
      owner = k < cap_a ? 0 : 1;
      int kk = owner ? k - cap_a : k;
      if (sub == owner) {
#pragma unroll 1
        for (unsigned long long m = L.mask; m; m &= m - 1) {
          const unsigned c = w.mv[pair_entry_addr(L.j0 + __ffsll((long long)m) - 1, L.nc0)];
          if (w.sq[c & 0x7fu] != 0 && kk-- == 0) {
            mine = c;
            break;
          }
        }
      }
    }
  }
  if (owner < 0) {
    const int k = (int)(x[0] % (uint32_t)n);
    owner = k < L.n_a ? 0 : 1;
    const int kk = owner ? k - L.n_a : k;
    if (sub == owner) mine = w.mv[pair_entry_addr(L.j0 + nth_bit64(L.mask, kk), L.nc0)];
  }
  const unsigned theirs = Pair::other(mine);
  return (sub == owner ? mine : theirs) & ~(unsigned)kCandIllegal;
}

// shared pick rule: index into the pair's list; both lanes compute the same value
__device__ __forceinline__ int pair_pick(const ThreadBoard& w, int n, int n0, uint64_t seed,
                                         uint32_t game_id, uint32_t ply, int capture_bias) {
  uint32_t x[4];
  philox4x32(game_id, ply, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  if (capture_bias > 0 && (int)(x[1] & 0xFFu) < capture_bias) {
    int ncap = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) ncap += w.sq[pair_move_at(w, i, n0) & 0x7fu] != 0;
    if (ncap > 0) {
      int k = (int)(x[0] % (uint32_t)ncap);
#pragma unroll 1
      for (int i = 0; i < n; ++i)
        if (w.sq[pair_move_at(w, i, n0) & 0x7fu] != 0 && k-- == 0) return i;
    }
  }
  return (int)(x[0] % (uint32_t)n);
}

}  // namespace xq
#endif  // __CUDACC__
