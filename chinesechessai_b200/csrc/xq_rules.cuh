// xq_rules.cuh — warp-per-board Xiangqi rules engine for sm_100a.
//
// One warp owns one board.  The 90 squares, per-row / per-column occupancy
// masks, the own-piece list, the pseudo-legal candidate list and the legal
// move list live in a 1 KB shared-memory slab per warp; game scalars are
// warp-uniform registers.  Semantics restate the reference's rules engine
// (chess_env.py; file:line cited per function), quirks included (SURVEY.md
// Appendix A).  Unlike the reference, which simulates each candidate on a
// board copy and regenerates every opposing piece's moves (:431-464,:506-548),
// legality is decided by an inverse, king-outward attack test on occupancy
// bit masks with the candidate applied as an override — no board copies.
#pragma once

#include <stdint.h>

#include "../../include/xq_b200.h"

// Per-lane logic (attack test, suicide filter, candidate generation, reward
// arithmetic) is XQ_HD so that tests/host_mirror can compile exactly this code
// with g++ and fuzz it on the CPU; warp-collective code is CUDA-only.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define XQ_HD __host__ __device__ __forceinline__
#define XQ_ALIGN16 __align__(16)
#else
#include <algorithm>
#include <cstdlib>
#define XQ_HD inline
#define XQ_ALIGN16 alignas(16)
#endif

namespace xq {

XQ_HD int xq_ffs(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}
XQ_HD int xq_clz(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
XQ_HD unsigned xq_brev(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
#endif
}
XQ_HD double xq_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
XQ_HD double xq_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
XQ_HD int xq_min(int a, int b) { return a < b ? a : b; }
XQ_HD int xq_max(int a, int b) { return a > b ? a : b; }
XQ_HD int xq_abs(int a) { return a < 0 ? -a : a; }

constexpr int kWarpsPerCta = 8;
constexpr unsigned kFull = 0xffffffffu;

enum : int { KING = 1, ADVISOR = 2, BISHOP = 3, KNIGHT = 4, ROOK = 5, CANNON = 6, PAWN = 7 };

// Per-warp shared-memory slab (1536 B).
struct XQ_ALIGN16 WarpSmem {
  int8_t sq[XQ_BOARD_STRIDE];  // board, row-major r*9+c (chess_env.py:17)
  uint16_t rows[16];           // rows[r] bit c = square (r,c) occupied
  uint16_t cols[16];           // cols[c] bit r = square (r,c) occupied
  uint8_t own[96];             // squares of the side-to-move's pieces, row-major order
  // candidates in canonical order: from<<8 | to; bit 7 = "cannot affect king safety"
  // (legal iff the position itself is safe), bit 15 = tested and found illegal
  uint16_t cand[XQ_CAND_CAP];
  uint16_t wl[XQ_CAND_CAP];    // worklist of candidate indices that need the full test
  int16_t moves[XQ_MAX_MOVES]; // legal moves, packed from*90+to
};
static_assert(sizeof(WarpSmem) == 1536, "WarpSmem must be 1.5 KB");
constexpr uint16_t kCandIrrelevant = 0x0080, kCandIllegal = 0x8000, kWlSentinel = 0xFFFF;

// Warp-uniform game scalars (chess_env.py:17-31,62-65).
struct Game {
  int player, winner, reason, done, red_king, black_king, flags;
  int move_count, no_capture, cchecks, hist_len, check_len;
  unsigned check_bits;
  uint64_t bkey;  // position key of the staged board without the side byte (kept incrementally)
};


XQ_HD uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

#if defined(__CUDACC__)
// A board (or tree) is owned by a TILE of L consecutive lanes, L in {8,16,32}: a warp carries
// 32/L boards.  The per-board scalar work (make_move bookkeeping, pick, digest, loop control)
// is then shared by 32/L boards per issued instruction.  Every collective uses the tile's own
// member mask, so tiles of one warp may diverge freely.
template <int L>
struct Tile {
  static_assert(L == 8 || L == 16 || L == 32, "tile width");
  static __device__ __forceinline__ int lane() { return threadIdx.x & (L - 1); }
  static __device__ __forceinline__ int shift() { return (threadIdx.x & 31) & ~(L - 1); }
  static __device__ __forceinline__ unsigned lanes() { return L == 32 ? 0xffffffffu : ((1u << (L & 31)) - 1u); }
  static __device__ __forceinline__ unsigned mask() { return L == 32 ? 0xffffffffu : lanes() << shift(); }
  static __device__ __forceinline__ unsigned ballot(bool p) {
    const unsigned b = __ballot_sync(mask(), p);
    return L == 32 ? b : (b >> shift()) & lanes();
  }
  static __device__ __forceinline__ bool any(bool p) { return ballot(p) != 0; }
  template <typename T>
  static __device__ __forceinline__ T shfl(T v, int src) { return __shfl_sync(mask(), v, src, L); }
  template <typename T>
  static __device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(mask(), v, d, L); }
  template <typename T>
  static __device__ __forceinline__ T shfl_xor(T v, int d) { return __shfl_xor_sync(mask(), v, d, L); }
  static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
  static __device__ __forceinline__ unsigned sum(unsigned v) { return __reduce_add_sync(mask(), v); }
  static __device__ __forceinline__ uint64_t xor64(uint64_t v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v ^= shfl_xor(v, o);
    return v;
  }
};

#endif  // __CUDACC__

XQ_HD uint64_t side_key(int player) {
  return mix64(0x7000ULL + (player == 1 ? 0u : 1u));  // chess_env.py:503
}
XQ_HD uint64_t piece_key(int piece, int s) { return mix64((uint64_t)((piece + 8) * 128 + s)); }

#if defined(__CUDACC__)

// Position key without the side byte: XOR of per-(piece,square) keys, computed
// lane-parallel.  _get_position_hash, chess_env.py:497-504.
template <int L>
__device__ __forceinline__ uint64_t board_key(const WarpSmem& w) {
  uint64_t h = 0;
#pragma unroll 1
  for (int s = Tile<L>::lane(); s < XQ_NSQ; s += L) {
    const int p = w.sq[s];
    if (p != 0) h ^= piece_key(p, s);
  }
  return Tile<L>::xor64(h);
}

// ---- state load / store ---------------------------------------------------
template <int L>
__device__ __forceinline__ void load_board(WarpSmem& w, const int8_t* __restrict__ board_row) {
#pragma unroll
  for (int i = Tile<L>::lane(); i < XQ_BOARD_STRIDE / 4; i += L)
    reinterpret_cast<uint32_t*>(w.sq)[i] = reinterpret_cast<const uint32_t*>(board_row)[i];
  Tile<L>::sync();
}

template <int L>
__device__ __forceinline__ void store_board(const WarpSmem& w, int8_t* __restrict__ board_row) {
  Tile<L>::sync();
#pragma unroll
  for (int i = Tile<L>::lane(); i < XQ_BOARD_STRIDE / 4; i += L)
    reinterpret_cast<uint32_t*>(board_row)[i] = reinterpret_cast<const uint32_t*>(w.sq)[i];
}

__device__ __forceinline__ Game load_meta(const xq_meta* __restrict__ m) {
  const uint4* p = reinterpret_cast<const uint4*>(m);
  uint4 a = p[0], b = p[1];
  Game g;
  g.player = (int8_t)(a.x & 0xff);
  g.winner = (int8_t)((a.x >> 8) & 0xff);
  g.reason = (a.x >> 16) & 0xff;
  g.done = (a.x >> 24) & 0xff;
  g.red_king = (int8_t)(a.y & 0xff);
  g.black_king = (int8_t)((a.y >> 8) & 0xff);
  g.flags = (a.y >> 16) & 0xff;
  g.move_count = (int)a.z;
  g.no_capture = (int)a.w;
  g.cchecks = (int)b.x;
  g.hist_len = (int)b.y;
  g.check_bits = b.z;
  g.check_len = (int)b.w;
  return g;
}

template <int L>
__device__ __forceinline__ void store_meta(xq_meta* __restrict__ m, const Game& g) {
  if (Tile<L>::lane() == 0) {
    uint4 a, b;
    a.x = (uint32_t)(g.player & 0xff) | ((uint32_t)(g.winner & 0xff) << 8) |
          ((uint32_t)(g.reason & 0xff) << 16) | ((uint32_t)(g.done & 0xff) << 24);
    a.y = (uint32_t)(g.red_king & 0xff) | ((uint32_t)(g.black_king & 0xff) << 8) |
          ((uint32_t)(g.flags & 0xff) << 16);
    a.z = (uint32_t)g.move_count;
    a.w = (uint32_t)g.no_capture;
    b.x = (uint32_t)g.cchecks;
    b.y = (uint32_t)g.hist_len;
    b.z = g.check_bits;
    b.w = (uint32_t)g.check_len;
    uint4* p = reinterpret_cast<uint4*>(m);
    p[0] = a;
    p[1] = b;
  }
}

// Row / column occupancy masks from the staged board (19 tasks over the tile's lanes).
template <int L>
__device__ __forceinline__ void build_masks(WarpSmem& w) {
#pragma unroll 1
  for (int t = Tile<L>::lane(); t < 19; t += L) {
    unsigned m = 0;
    if (t < 10) {
#pragma unroll
      for (int c = 0; c < 9; ++c) m |= (w.sq[t * 9 + c] != 0 ? 1u : 0u) << c;
      w.rows[t] = (uint16_t)m;
    } else {
      const int c = t - 10;
#pragma unroll
      for (int r = 0; r < 10; ++r) m |= (w.sq[r * 9 + c] != 0 ? 1u : 0u) << r;
      w.cols[c] = (uint16_t)m;
    }
  }
  Tile<L>::sync();
}

#endif  // __CUDACC__

// ---- attack test -----------------------------------------------------------
// Is square K a pseudo-target of any piece of sign `es` (the attackers), with
// K/A/B/P geometry taken from side `geo` (chess_env.py:506-548 — the reference
// regenerates attackers' moves with self.current_player's geometry, quirk A.3)?
// A candidate (from,to,mover) can be applied as an override (from<0: none).
// The test is 8 probes from K outward: 4 rays (first piece = rook / adjacent pawn / adjacent
// king; second piece = cannon) and 4 diagonal neighbours (knight legs :182-197, bishop eyes
// :161-174, advisors :149-152).  Probes are written once and looped (code size matters: the
// fused playout loop must stay inside the 32 KB L1.5 instruction cache).
struct Probe {
  int K, kr, kc, es, geo, from, to, mover;
  unsigned rowm, colm;  // occupancy of K's row / column with the override applied
  bool kocc, in_pal;
};

template <class W>
XQ_HD Probe make_probe(const W& w, int K, int es, int geo, int from, int to, int mover) {
  Probe p;
  p.K = K; p.es = es; p.geo = geo; p.from = from; p.to = to; p.mover = mover;
  p.kr = K / 9;
  p.kc = K - p.kr * 9;
  p.rowm = w.rows[p.kr];
  p.colm = w.cols[p.kc];
  if (from >= 0) {
    const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
    if (fr == p.kr) p.rowm &= ~(1u << fc);
    if (fc == p.kc) p.colm &= ~(1u << fr);
    if (tr == p.kr) p.rowm |= 1u << tc;
    if (tc == p.kc) p.colm |= 1u << tr;
  }
  p.kocc = (p.rowm >> p.kc) & 1u;
  p.in_pal = (p.kc >= 3 && p.kc <= 5) && (geo == 1 ? p.kr >= 7 : p.kr <= 2);  // :127-131
  return p;
}

template <class W>
XQ_HD int probe_piece(const W& w, const Probe& p, int s) {
  return s == p.to ? p.mover : (s == p.from ? 0 : (int)w.sq[s]);
}

// dir 0..3 = rays (0,+1),(0,-1),(+1,0),(-1,0).  The mask is mirrored for the backward rays so
// "ahead" is always toward higher bits.
template <class W>
XQ_HD bool probe_ray(const W& w, const Probe& p, int dir) {
  const bool horiz = dir < 2, fwd = (dir & 1) == 0;
  const int len = horiz ? 9 : 10;
  unsigned m = horiz ? p.rowm : p.colm;
  int x = horiz ? p.kc : p.kr;
  if (!fwd) {
    m = xq_brev(m) >> (32 - len);
    x = len - 1 - x;
  }
  const unsigned a = m >> (x + 1);
  if (!a) return false;
  const int delta = (horiz ? 1 : 9) * (fwd ? 1 : -1);
  const int d1 = xq_ffs(a), q = probe_piece(w, p, p.K + d1 * delta);
  // adjacent pawn: sideways only from a crossed row (:242,:247); vertically only if it moves
  // toward K under `geo` (a pawn below K attacks iff pawns move to smaller rows, geo==1, :241)
  const bool pawn_dir = horiz ? (p.geo == 1 ? p.kr < 5 : p.kr >= 5) : (p.geo == (fwd ? 1 : -1));
  bool hit = (q == p.es * ROOK) | (q == p.es * CANNON && !p.kocc) |
             (d1 == 1 && ((q == p.es * PAWN && pawn_dir) | (q == p.es * KING && p.in_pal)));
  const unsigned a2 = a & (a - 1);
  if (p.kocc && a2) hit |= probe_piece(w, p, p.K + xq_ffs(a2) * delta) == p.es * CANNON;
  return hit;
}

// i 0..3 = diagonal neighbour (a,b) in {-1,+1}^2.
template <class W>
XQ_HD bool probe_diag(const W& w, const Probe& p, int i, bool exotic) {
  const int a = (i & 2) ? 1 : -1, b = (i & 1) ? 1 : -1;
  const int lr = p.kr + a, lc = p.kc + b;
  if (lr < 0 || lr > 9 || lc < 0 || lc > 8) return false;
  const int ql = probe_piece(w, p, lr * 9 + lc);
  if (ql != 0) return exotic && p.in_pal && ql == p.es * ADVISOR;
  const int r2 = p.kr + 2 * a, c2 = p.kc + 2 * b;
  const bool r2ok = r2 >= 0 && r2 <= 9, c2ok = c2 >= 0 && c2 <= 8;
  const bool bside = p.geo == 1 ? p.kr >= 5 : p.kr <= 3;  // :159,:167-170 (black river = 4)
  bool hit = false;
  if (r2ok) hit |= probe_piece(w, p, r2 * 9 + lc) == p.es * KNIGHT;
  if (c2ok) hit |= probe_piece(w, p, lr * 9 + c2) == p.es * KNIGHT;
  if (exotic && bside && r2ok && c2ok) hit |= probe_piece(w, p, r2 * 9 + c2) == p.es * BISHOP;
  return hit;
}

// `exotic` enables the K/A/B diagonal probes (warp-uniform hint; always safe to pass true).
template <class W>
XQ_HD bool attacked(const W& w, int K, int es, int geo, int from, int to, int mover,
                    bool exotic, unsigned* colm_out) {
  const Probe p = make_probe(w, K, es, geo, from, to, mover);
  if (colm_out) *colm_out = p.colm;
  bool hit = false;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int d = 0; d < 4; ++d) hit |= probe_ray(w, p, d);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 4; ++i) hit |= probe_diag(w, p, i, exotic);
  return hit;
}

// One of the 8 probes of attacked() on the staged board without an override (dir 4..7 = the
// diagonals); the probe round of movegen<L> runs these on 8 lanes.  Kept as a function for the
// host mirror, which checks that the OR over the 8 directions equals attacked().
template <class W>
XQ_HD bool attacked_dir(const W& w, int K, int es, int geo, int dir) {
  const Probe p = make_probe(w, K, es, geo, -1, -1, 0);
  return dir < 4 ? probe_ray(w, p, dir) : probe_diag(w, p, dir - 4, true);
}

// _is_in_check(player) on the staged board (chess_env.py:506-548), single-lane form.
template <class W>
XQ_HD bool in_check(const W& w, const Game& g, int player) {
  const int K = player == 1 ? g.red_king : g.black_king;
  if (K < 0) return false;  // :517
  return attacked(w, K, -player, g.player, -1, -1, 0, true, nullptr);
}

// _is_move_suicide (chess_env.py:431-464): own king attacked after the move
// (geometry of the side to move) OR cached kings face each other (:466-495;
// only the MOVING king's cache is refreshed, :448-451 — stale-cache quirk A.4).
// from < 0 evaluates the position itself (no move).
template <class W>
XQ_HD bool suicide(const W& w, const Game& g, int from, int to, bool exotic) {
  const int mover = from >= 0 ? (int)w.sq[from] : 0;
  int red = g.red_king, black = g.black_king;
  if (mover == KING) red = to;
  else if (mover == -KING) black = to;
  const int K = g.player == 1 ? red : black;
  unsigned colm = 0;
  bool bad = false;
  if (K >= 0) bad = attacked(w, K, -g.player, g.player, from, to, mover, exotic, &colm);
  if (red >= 0 && black >= 0) {
    const int rr = red / 9, rc = red - rr * 9, br = black / 9, bc = black - br * 9;
    if (rc == bc) {
      const int lo = xq_min(rr, br), hi = xq_max(rr, br);
      const unsigned between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
      bad |= (colm & between) == 0;
    }
  }
  return bad;
}

// Squares whose content attacked()/facing can read for a king on (kr,kc) when the K/A/B
// probes are off: its row, its column, the 4 diagonal neighbours and the 8 knight squares.
// A non-king move touching none of them leaves the king exactly as safe as it is now.
XQ_HD bool touches(int kr, int kc, int s) {
  const int r = s / 9, c = s - r * 9;
  const int dr = xq_abs(r - kr), dc = xq_abs(c - kc);
  return dr == 0 || dc == 0 || (dr <= 2 && dc <= 2 && dr + dc <= 3);
}

// ---- fast legality test for regular positions -----------------------------------
// Regular = exactly one own king on its cached square and no enemy K/A/B within 3 rows of it
// (every position reachable in play).  Then only R, C, N and P can attack the king, the king
// square is occupied, and for a NON-king move the test reduces to bit operations on masks that
// are computed once per position: occupancy of the king's row/column, the enemy rooks / cannons
// / pawns on them, the enemy knights on the 8 knight squares and the 4 leg squares.
struct FastCtx {
  int K, kr, kc, geo;
  unsigned rowm, colm;              // occupancy of K's row (bit = column) / column (bit = row)
  unsigned er_row, ec_row, ep_row;  // enemy rooks / cannons / pawns on K's row
  unsigned er_col, ec_col, ep_col;  // ... on K's column
  unsigned ekn;                     // bit diag+4t: enemy knight on K+(2a,b) [t=0] / K+(a,2b) [t=1]
  unsigned legocc;                  // bit diag: leg K+(a,b) occupied or off-board
  // derived once per position (finish_fast_ctx): the test itself is then mask algebra only
  unsigned row_lo, row_hi;          // the bits below / above the king in its row
  unsigned col_lo, col_hi;          // ... in its column
  unsigned prow, pcol;              // squares of the row / column an enemy pawn attacks K from
  unsigned facing;                  // rows strictly between the two kings on a shared file; all ones
                                    // when they are on different files (the test can then never fire:
                                    // the king's own bit is in colm)
  const uint32_t* touch;            // row K of the touch table (xq_touch_table.inc)
};

// Touch table (generated, xq_touch_table.inc): entry [K][s] = what a piece leaving or landing on
// square s changes in the masks above for a king on K — the bit of s in the king's row (bits 0-8)
// and column (9-18), the leg bit if s is a diagonal neighbour (19-22), the knight bit if s is one
// of the eight knight squares (23-30).  32 KB; a game's kings visit a handful of rows of it.
constexpr int kTouchEntries = 90 * 90;
XQ_HD uint32_t touch_entry(const uint32_t* row, int s) {
#if defined(__CUDA_ARCH__)
  return __ldg(row + s);
#else
  return row[s];
#endif
}

// diag index of probe_diag: bit1 = (dr>0), bit0 = (dc>0)
XQ_HD int diag_index(int dr, int dc) { return (dr > 0 ? 2 : 0) + (dc > 0 ? 1 : 0); }

// Pawn-attack squares for geometry `geo` (chess_env.py:240-249 as seen from the attacked king):
// sideways from the adjacent columns once the pawn has crossed the river (the king then stands
// on the far side for `geo`), vertically from the one adjacent row the pawn moves away from.
XQ_HD void set_pawn_masks(FastCtx& f, int geo) {
  const bool side_ok = geo == 1 ? f.kr < 5 : f.kr >= 5;
  f.prow = side_ok ? (((2u << f.kc) | ((1u << f.kc) >> 1)) & 0x1FFu) : 0u;
  f.pcol = geo == 1 ? ((2u << f.kr) & 0x3FFu) : ((1u << f.kr) >> 1);
}
// The fields derived from K, the geometry and the enemy king's cached square.
XQ_HD void finish_fast_ctx(FastCtx& f, int player, int enemy_king) {
  f.row_lo = (1u << f.kc) - 1u;
  f.row_hi = ~((2u << f.kc) - 1u) & 0x1FFu;
  f.col_lo = (1u << f.kr) - 1u;
  f.col_hi = ~((2u << f.kr) - 1u) & 0x3FFu;
  set_pawn_masks(f, player);
  f.facing = 0xFFFFFFFFu;
  if (enemy_king >= 0 && enemy_king % 9 == f.kc) {
    const int er = enemy_king / 9, lo = xq_min(er, f.kr), hi = xq_max(er, f.kr);
    f.facing = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  }
}

// Sequential construction (host mirror); the kernel builds the same masks with ballots.
template <class W>
XQ_HD FastCtx make_fast_ctx(const W& w, const Game& g, const uint32_t* touch_table) {
  FastCtx f{};
  const int player = g.player, es = -player;
  f.K = player == 1 ? g.red_king : g.black_king;
  f.touch = touch_table + f.K * 90;
  f.kr = f.K / 9;
  f.kc = f.K - f.kr * 9;
  f.geo = player;
  f.rowm = w.rows[f.kr];
  f.colm = w.cols[f.kc];
  for (int c = 0; c < 9; ++c) {
    const int q = w.sq[f.kr * 9 + c];
    f.er_row |= (q == es * ROOK ? 1u : 0u) << c;
    f.ec_row |= (q == es * CANNON ? 1u : 0u) << c;
    f.ep_row |= (q == es * PAWN ? 1u : 0u) << c;
  }
  for (int r = 0; r < 10; ++r) {
    const int q = w.sq[r * 9 + f.kc];
    f.er_col |= (q == es * ROOK ? 1u : 0u) << r;
    f.ec_col |= (q == es * CANNON ? 1u : 0u) << r;
    f.ep_col |= (q == es * PAWN ? 1u : 0u) << r;
  }
  for (int d = 0; d < 4; ++d) {
    const int a = (d & 2) ? 1 : -1, b = (d & 1) ? 1 : -1;
    const int lr = f.kr + a, lc = f.kc + b;
    const bool lon = lr >= 0 && lr <= 9 && lc >= 0 && lc <= 8;
    if (!lon || w.sq[lr * 9 + lc] != 0) f.legocc |= 1u << d;
    const int r2 = f.kr + 2 * a, c2 = f.kc + 2 * b;
    if (lon && r2 >= 0 && r2 <= 9 && w.sq[r2 * 9 + lc] == es * KNIGHT) f.ekn |= 1u << d;
    if (lon && c2 >= 0 && c2 <= 8 && w.sq[lr * 9 + c2] == es * KNIGHT) f.ekn |= 1u << (4 + d);
  }
  finish_fast_ctx(f, player, player == 1 ? g.black_king : g.red_king);
  return f;
}

// Kings-facing verdict for a move of the own king to `to` (chess_env.py:448-451,:466-495);
// p = probe built at the king's NEW square with the move applied (the probe round inlines the
// same test; this form is what the host mirror exercises).
XQ_HD bool king_move_facing(const Game& g, const Probe& p, int to) {
  const int ek = g.player == 1 ? g.black_king : g.red_king;
  if (ek < 0) return false;
  const int er = ek / 9, ec = ek - er * 9;
  if (ec != p.kc) return false;
  const int lo = xq_min(er, p.kr), hi = xq_max(er, p.kr);
  const unsigned between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  (void)to;
  return (p.colm & between) == 0;
}

// isolate the lowest / highest set bit (0 if none)
XQ_HD unsigned low_bit(unsigned x) { return x & (0u - x); }
XQ_HD unsigned high_bit(unsigned x) {
#if defined(__CUDA_ARCH__)
  // bfind gives the index of the leading one, or 0xFFFFFFFF for 0; shl clamps shift amounts
  // above 31 to 32, i.e. to a zero result: two instructions, no compare / select
  unsigned idx, r;
  asm("bfind.u32 %0, %1;" : "=r"(idx) : "r"(x));
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(1u), "r"(idx));
  return r;
#else
  return x ? 0x80000000u >> xq_clz(x) : 0u;
#endif
}

// _is_move_suicide for a non-king move of the side to move in a regular position
// (from < 0: the position itself).  Same verdict as suicide(); pure register arithmetic.
XQ_HD bool suicide_fast(const FastCtx& f, int from, int to) {
  unsigned rowm = f.rowm, colm = f.colm;
  unsigned er_row = f.er_row, ec_row = f.ec_row, ep_row = f.ep_row;
  unsigned er_col = f.er_col, ec_col = f.ec_col, ep_col = f.ep_col;
  unsigned ekn = f.ekn, legocc = f.legocc;
  if (from >= 0) {
    // what leaving `from` and landing on `to` change, from the touch table (no divisions, no
    // compares): the square's bit in the king's row / column, its leg bit, its knight bit
    const uint32_t mf = touch_entry(f.touch, from), mt = touch_entry(f.touch, to);
    const unsigned rt = mt & 0x1FFu, ct = (mt >> 9) & 0x3FFu;
    rowm = (rowm & ~(mf & 0x1FFu)) | rt;        // own piece lands on the row: occupied,
    er_row &= ~rt; ec_row &= ~rt; ep_row &= ~rt;  // whatever stood there is gone
    colm = (colm & ~((mf >> 9) & 0x3FFu)) | ct;
    er_col &= ~ct; ec_col &= ~ct; ep_col &= ~ct;
    legocc = (legocc & ~(mf >> 19)) | ((mt >> 19) & 0xFu);
    ekn &= ~(mt >> 23);                         // lands on a knight square: that knight is gone
  }
  // nearest and second-nearest piece on each of the four rays, as bits of the row / column
  const unsigned ra = rowm & f.row_hi, rb = rowm & f.row_lo, ca = colm & f.col_hi, cb = colm & f.col_lo;
  const unsigned f1 = low_bit(ra), f2 = high_bit(rb), f3 = low_bit(ca), f4 = high_bit(cb);
  const unsigned s1 = low_bit(ra ^ f1), s2 = high_bit(rb ^ f2), s3 = low_bit(ca ^ f3), s4 = high_bit(cb ^ f4);
  // rook = nearest piece, cannon = second piece (the nearest is its screen), pawn = adjacent
  // square on the side it attacks from; knights: a free leg exposes the two squares behind it
  unsigned h = ((f1 | f2) & er_row) | ((s1 | s2) & ec_row) | (ep_row & f.prow) |
               ((f3 | f4) & er_col) | ((s3 | s4) & ec_col) | (ep_col & f.pcol) |
               ((ekn | (ekn >> 4)) & ~legocc & 0xFu);
  // kings facing (:466-495): caches are unchanged by a non-king move
  return h != 0 || (colm & f.facing) == 0;
}

// make_move's check test (chess_env.py:317: is the side now to move in check, geometry of the
// side that just moved) for a REGULAR position, from the masks of the legality test: flip the
// pawn geometry to the attackers' own (they are the previous mover's pieces, so here the
// geometry is the correct one, quirk A.3) and drop the kings-facing term, which is not part of
// _is_in_check.  Regular also means no enemy K/A/B within reach of the king's row, so R, C, N, P
// are the only possible attackers, exactly the ones the masks hold.
XQ_HD bool check_fast(const FastCtx& f, int player) {
  FastCtx c = f;
  set_pawn_masks(c, -player);  // the attackers' own geometry
  c.facing = 0xFFFFFFFFu;
  return suicide_fast(c, -1, -1);
}

// ---- move generation --------------------------------------------------------
// One work item of candidate generation: (own piece at `from`, direction d of
// the reference's per-piece generator order).  Produces, in generator order,
// `empties` quiet ray steps from+k*delta (rook/cannon only) followed by up to
// two explicit targets e1, e2 (-1 = none).  Already filtered by on-board (:113)
// and not-own-piece (:116).
struct Item {
  int from, empties, delta, e1, e2;
};

// Leaper table (generated, xq_leap_table.inc): entry [side][piece type][from][d] =
// t1 | t2<<8 | blocker<<16: the targets of generator slot d (a square, or the piece's own square
// for "none") that are on-board and inside the zone the generator enforces (palace :127-147, own river side
// :159-170, pawn direction/crossing :240-249) and the square that must be empty (knight leg
// :189-195, bishop eye :171-174).  23 KB, read through the L1/read-only path.
constexpr int kLeapEntries = 2 * 8 * 90 * 4;
XQ_HD int leap_index(int player, int pt, int from, int d) {
  return ((player == 1 ? 0 : 8) + pt) * 360 + from * 4 + d;
}

// Ray table (generated, xq_ray_table.inc): entry [orientation][position][occupancy mask][backward]
// = empties | first << 4 | second << 8 for a rook / cannon ray (see gen_ray_table.py).  58 KB.
constexpr int kRayRowEntries = 9 * 512 * 2, kRayEntries = kRayRowEntries + 10 * 1024 * 2;
// Knight table (generated, xq_knight_table.inc): entry [T][diagonal] = leg | knight_a << 8 |
// knight_b << 16, the squares from which a knight attacks T (0xFF = off the board).  1.4 KB.
constexpr int kKnightEntries = 90 * 4;
struct Tables {
  const uint32_t* leap;
  const uint16_t* ray;
  const uint32_t* knight = nullptr;  // only king_move_fast() reads it
};

// One generator slot of one piece, branch-free: the lanes of a warp work on different piece types
// at the same loop index, so a ray path and a leaper path would be issued one after the other
// (measured: 15.6 of 32 lanes inside this function).  Both kinds reduce to the same shape —
// one table word, a square that must be empty, two target squares with the own-piece test
// (:116), a run of quiet squares — with "no target" named by the piece's own square.
template <class W>
XQ_HD Item gen_item(const W& w, const Tables& tb, int player, int from, int d) {
  const int p = w.sq[from];
  const int pt = p < 0 ? -p : p;
  const bool ray = pt == ROOK || pt == CANNON;  // :199-235, rays (0,1),(0,-1),(1,0),(-1,0)
  const bool leaper = !ray && pt >= KING && pt <= PAWN;
  const int r = (from * 57) >> 9, c = from - r * 9;  // from / 9 for from < 90
  const bool horiz = d < 2;
  const int back = d & 1;
  const int delta = (horiz ? 1 : 9) * (back ? -1 : 1);
  // the occupancy mask of the piece's row or column: rows[] and cols[] are adjacent members
  const int line = (int)(&w.rows[0])[horiz ? r : (int)(&w.cols[0] - &w.rows[0]) + c];
  const int ridx = horiz ? ((c * 512 + line) * 2 + back) : (kRayRowEntries + (r * 1024 + line) * 2 + back);
  const int lidx = leap_index(player, pt & 7, from, d);
  // a piece code outside 1..7 (poked boards) generates nothing: both targets = own square
  unsigned e = (unsigned)from * 0x101u | 0xFF0000u;
#if defined(__CUDA_ARCH__)
  if (ray) e = __ldg(tb.ray + ridx);
  if (leaper) e = __ldg(tb.leap + lidx);
#else
  if (ray) e = tb.ray[ridx];
  if (leaper) e = tb.leap[lidx];
#endif
  const int hitd = (int)((pt == CANNON ? e >> 8 : e >> 4) & 15u);  // 0: nothing to capture
  const int t1 = ray ? from + hitd * delta : (int)(e & 0xFFu);
  const int t2 = ray ? from : (int)((e >> 8) & 0xFFu);
  const unsigned bk = ray ? 0xFFu : (e >> 16) & 0xFFu;  // knight leg :189-195, bishop eye :171-174
  const bool open = bk == 0xFFu || w.sq[bk == 0xFFu ? from : (int)bk] == 0;
  Item it;
  it.from = from;
  it.delta = delta;
  it.empties = ray ? (int)(e & 15u) : 0;
  it.e1 = (open && (int)w.sq[t1] * player <= 0) ? t1 : -1;  // :116
  it.e2 = (open && (int)w.sq[t2] * player <= 0) ? t2 : -1;
  return it;
}

// The same generator, one PIECE at a time (the lane-pair engine): what gen_item() derives per slot
// — piece type, row / column, table address — is derived once, the four table words of the piece
// come in one 16-byte load (leapers: the four slots are adjacent) or two 4-byte loads (rays: the
// forward and backward entries of a line are adjacent), and the four slots are decoded by
// straight-line code with compile-time directions.
struct PieceGen {
  uint32_t e[4];     // table word per slot, "no target" = own square (see gen_item)
  int from;
  unsigned cap_sh;   // bit offset of the capture distance in a ray word (cannon 8, rook 4)
  bool ray;
};
template <class W>
XQ_HD PieceGen gen_piece(const W& w, const Tables& tb, int player, int from) {
  PieceGen g;
  const int p = w.sq[from];
  const int pt = p < 0 ? -p : p;
  g.from = from;
  g.ray = (unsigned)(pt - ROOK) < 2u;
  g.cap_sh = pt == CANNON ? 8u : 4u;
  const uint32_t none = (unsigned)from * 0x101u | 0xFF0000u;
  g.e[0] = g.e[1] = g.e[2] = g.e[3] = none;
  if (g.ray) {
    const int r = (from * 57) >> 9, c = from - r * 9;
    const uint32_t* ray32 = reinterpret_cast<const uint32_t*>(tb.ray);
    const int hi = c * 512 + (int)w.rows[r], vi = kRayRowEntries / 2 + r * 1024 + (int)w.cols[c];
#if defined(__CUDA_ARCH__)
    const uint32_t h = __ldg(ray32 + hi), v = __ldg(ray32 + vi);
#else
    const uint32_t h = ray32[hi], v = ray32[vi];
#endif
    g.e[0] = h & 0xFFFFu;
    g.e[1] = h >> 16;
    g.e[2] = v & 0xFFFFu;
    g.e[3] = v >> 16;
  } else if ((unsigned)(pt - KING) <= (unsigned)(PAWN - KING)) {
    const uint32_t* q = tb.leap + leap_index(player, pt, from, 0);
#if defined(__CUDA_ARCH__)
    const uint4 L = __ldg(reinterpret_cast<const uint4*>(q));
    g.e[0] = L.x;
    g.e[1] = L.y;
    g.e[2] = L.z;
    g.e[3] = L.w;
#else
    for (int d = 0; d < 4; ++d) g.e[d] = q[d];
#endif
  }
  return g;
}
// slot d of the piece (d is a compile-time constant at the call sites)
template <class W>
XQ_HD Item gen_dir(const W& w, const PieceGen& g, int player, int d) {
  const uint32_t e = g.e[d];
  const int from = g.from;
  const int delta = (d < 2 ? 1 : 9) * ((d & 1) ? -1 : 1);
  const int hitd = (int)((e >> g.cap_sh) & 15u);  // 0: nothing to capture
  const int t1 = g.ray ? from + hitd * delta : (int)(e & 0xFFu);
  const int t2 = g.ray ? from : (int)((e >> 8) & 0xFFu);
  const unsigned bk = g.ray ? 0xFFu : (e >> 16) & 0xFFu;
  const bool open = bk == 0xFFu || w.sq[bk == 0xFFu ? from : (int)bk] == 0;
  Item it;
  it.from = from;
  it.delta = delta;
  it.empties = g.ray ? (int)(e & 15u) : 0;
  it.e1 = (open && (int)w.sq[t1] * player <= 0) ? t1 : -1;  // :116
  it.e2 = (open && (int)w.sq[t2] * player <= 0) ? t2 : -1;
  return it;
}

// _is_move_suicide (chess_env.py:431-464) for the king's OWN move K -> T on a regular position
// (exotic_window: no enemy K/A/B can reach the squares the king can step to), as straight-line
// table code instead of the eight probes of attacked(): the two ray words of T's row and column
// (with K vacated) name the nearest and second piece in the four directions — rook / adjacent pawn
// under the mover's geometry (quirk A.3) / cannon behind its screen —, the knight table names the
// four legs and eight knight squares, and the kings-facing test (:466-495) uses the refreshed
// cache of the moving king (:448-451) against the enemy's cached square `ek`.
template <class W>
XQ_HD bool king_move_fast(const W& w, const Tables& tb, int player, int K, int T, int ek) {
  const int es = -player;
  const int kr = (K * 57) >> 9, kc = K - kr * 9, tr = (T * 57) >> 9, tc = T - tr * 9;
  unsigned rowm = w.rows[tr], colm = w.cols[tc];
  rowm &= ~(kr == tr ? 1u << kc : 0u);  // the king has left K (the tables ignore T's own bit)
  colm &= ~(kc == tc ? 1u << kr : 0u);
  const uint32_t* ray32 = reinterpret_cast<const uint32_t*>(tb.ray);
  const int hi = tc * 512 + (int)rowm, vi = kRayRowEntries / 2 + tr * 1024 + (int)colm;
#if defined(__CUDA_ARCH__)
  const uint32_t h = __ldg(ray32 + hi), v = __ldg(ray32 + vi);
  const uint4 kn = __ldg(reinterpret_cast<const uint4*>(tb.knight) + T);
  const uint32_t kn4[4] = {kn.x, kn.y, kn.z, kn.w};
#else
  const uint32_t h = ray32[hi], v = ray32[vi];
  const uint32_t* kn4 = tb.knight + T * 4;
#endif
  // pawns: sideways only from a crossed row, vertically only toward the king (:240-249 with the
  // mover's geometry: the defender's forward direction)
  const bool side_ok = player == 1 ? tr < 5 : tr >= 5;
  bool hit = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int d = 0; d < 4; ++d) {
    const uint32_t e = d == 0 ? h & 0xFFFFu : d == 1 ? h >> 16 : d == 2 ? v & 0xFFFFu : v >> 16;
    const int delta = (d < 2 ? 1 : 9) * ((d & 1) ? -1 : 1);
    const int first = (int)((e >> 4) & 15u), second = (int)((e >> 8) & 15u);
    const int q1 = w.sq[T + first * delta], q2 = w.sq[T + second * delta];
    const bool pawn_ok = d < 2 ? side_ok : player == (d == 2 ? 1 : -1);
    hit |= first != 0 && (q1 == es * ROOK || (first == 1 && pawn_ok && q1 == es * PAWN));
    hit |= second != 0 && q2 == es * CANNON;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) {
    // an off-board square reads K instead: the own king is no enemy knight, and K counts as an
    // empty leg anyway (the king has left it)
    const uint32_t e = kn4[i];
    const int leg = (int)(e & 0xFFu), a = (int)((e >> 8) & 0xFFu), b = (int)((e >> 16) & 0xFFu);
    const int legs = leg == 0xFF ? K : leg, as = a == 0xFF ? K : a, bs = b == 0xFF ? K : b;
    const bool free_leg = legs == K || w.sq[legs] == 0;
    hit |= free_leg && (w.sq[as] == es * KNIGHT || w.sq[bs] == es * KNIGHT);
  }
  if (ek >= 0) {
    const int er = (ek * 57) >> 9, ec = ek - er * 9;
    const int lo = xq_min(tr, er), hi_r = xq_max(tr, er);
    const unsigned between = ((1u << hi_r) - 1u) & ~((2u << lo) - 1u);
    hit |= ec == tc && (colm & between) == 0;
  }
  return hit;
}

// Warp-uniform hint for suicide(): can an enemy K/A/B ever matter?  They reach squares within
// two rows of themselves (bishop 2, advisor / king 1).  The squares that count in a regular
// position (exactly one own king piece, standing on its cached square) are the king's own square
// and the squares it can step to — one row away at most and, because _king_moves only generates
// targets inside the mover's palace (chess_env.py:127-131), never outside the palace rows.  An
// enemy K/A/B is therefore irrelevant unless it stands within two rows of those rows.  In play
// that never happens (K/A stay in their palace, B on its side of the river), so games from the
// start position never leave the fast path; round 1 used a flat 3-row window around the king,
// which a bishop on the river row next to an advanced king tripped in 0.4 % of all plies — and a
// pair on the slow path holds up its whole warp (38 % of the plies of the slowest groups).
// Poked boards (several kings, stale or missing cache, :448-451 re-creates the cache wherever a
// king piece moves) always take the full test.
XQ_HD void exotic_window(int player, int own_king, int* lo_sq, int* hi_sq) {
  const int kr = own_king / 9;
  const int pal_lo = player == 1 ? 7 : 0, pal_hi = player == 1 ? 9 : 2;
  int rmin = kr, rmax = kr;
  const int a = xq_max(kr - 1, pal_lo), b = xq_min(kr + 1, pal_hi);
  if (a <= b) {  // rows of the palace squares next to the king
    rmin = xq_min(rmin, a);
    rmax = xq_max(rmax, b);
  }
  *lo_sq = xq_max(rmin - 2, 0) * 9;
  *hi_sq = (xq_min(rmax + 2, 9) + 1) * 9;  // exclusive; always >= 3 rows wide
}
XQ_HD bool exotic_piece(int p, int s, int player, int own_king) {
  const int ap = p < 0 ? -p : p;
  int lo, hi;
  exotic_window(player, own_king, &lo, &hi);
  return (p * player < 0) && ap <= BISHOP && s >= lo && s < hi;
}
template <class W>
XQ_HD bool regular_king(const W& w, int player, int own_king, int n_own_kings) {
  return n_own_kings == 1 && own_king >= 0 && w.sq[own_king] == player * KING;
}

#if defined(__CUDACC__)
static __device__ __align__(16) const uint32_t g_leap[kLeapEntries] = {
#include "xq_leap_table.inc"
};
static __device__ const uint32_t g_touch[kTouchEntries] = {
#include "xq_touch_table.inc"
};
static __device__ __align__(16) const uint16_t g_ray[kRayEntries] = {
#include "xq_ray_table.inc"
};
static __device__ __align__(16) const uint32_t g_knight[kKnightEntries] = {
#include "xq_knight_table.inc"
};

// Cold paths are kept out of line: the fused loop's hot code has to stay small enough for the
// L1.5 instruction cache (profiles/r1: v2 72 KB -> no_instruction 3.0 stalls per issue).
static __device__ __noinline__ bool in_check_cold(const WarpSmem* wp, const Game* gp, int player) {
  return in_check(*wp, *gp, player);
}

// Irregular (poked) boards — several kings, stale or missing cache, enemy K/A/B near the
// king: every candidate through the general suicide() (:118).  Cold path.
static __device__ __noinline__ void legality_generic(WarpSmem* wp, const Game* gp, int ncand,
                                                     int lane, int stride) {
  WarpSmem& w = *wp;
  for (int j = lane; j < ncand; j += stride) {
    const int c = w.cand[j];
    if (suicide(w, *gp, c >> 8, c & 0x7f, true)) w.cand[j] = (uint16_t)(c | kCandIllegal);
  }
}

// get_legal_moves (chess_env.py:76-121) by one tile of L lanes.  Fills w.moves in the
// reference's order and returns the count.
// Phase A: work item = (own piece, direction) -> candidate list via an ordered scan.
// Phase B (regular boards): candidates that can change the king's safety go through the
// bitmask test suicide_fast(), the others inherit the verdict of the position itself; king
// moves and make_move's check test share one probe round.  Irregular boards: legality_generic.
// Ordered compaction with ballot/popc.
template <int L>
__device__ __forceinline__ int movegen(WarpSmem& w, Game& g, const uint32_t* __restrict__ leap,
                                       bool* checked_out = nullptr) {
  const Tables tb{leap, g_ray};
  using T = Tile<L>;
  const int lane = T::lane();
  const unsigned lt = (1u << lane) - 1u;
  const int player = g.player;

  // own-piece list in row-major order (:82-87) + "exotic" hint
  const int ownK = player == 1 ? g.red_king : g.black_king;
  int n_own = 0, n_kings = 0;
  bool ex = false;
#pragma unroll 1
  for (int base = 0; base < XQ_BOARD_STRIDE; base += L) {
    const int s = base + lane;
    const int p = s < XQ_NSQ ? (int)w.sq[s] : 0;
    const bool mine = p * player > 0;
    const unsigned b = T::ballot(mine);
    if (mine) w.own[n_own + __popc(b & lt)] = (uint8_t)s;
    n_own += __popc(b);
    n_kings += __popc(T::ballot(p == player * KING));
    ex |= exotic_piece(p, s, player, ownK < 0 ? 0 : ownK);
  }
  const bool exotic = !regular_king(w, player, ownK, n_kings) || T::any(ex);
  T::sync();

  // Phase A
  int ncand = 0;
  const int n_items = n_own * 4;
#pragma unroll 1
  for (int base = 0; base < n_items; base += L) {
    const int t = base + lane;
    Item it{0, 0, 0, -1, -1};
    if (t < n_items) it = gen_item(w, tb, player, w.own[t >> 2], t & 3);
    const int cnt = it.empties + (it.e1 >= 0) + (it.e2 >= 0);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
      const int v = T::shfl_up(incl, o);
      if (lane >= o) incl += v;
    }
    const int total = T::shfl(incl, L - 1);
    if (ncand + total > XQ_CAND_CAP) {  // tile-uniform; only on absurd poked boards
      g.flags |= XQ_F_OVERFLOW;
      break;
    }
    int off = ncand + incl - cnt;
    ncand += total;
    const int fs = it.from << 8;
#pragma unroll 1
    for (int k = 1; k <= it.empties; ++k, ++off) w.cand[off] = (uint16_t)(fs | (it.from + k * it.delta));
    if (it.e1 >= 0) w.cand[off++] = (uint16_t)(fs | it.e1);
    if (it.e2 >= 0) w.cand[off] = (uint16_t)(fs | it.e2);
  }
  T::sync();

  bool cur_bad = false;
  int kfirst = 0, kcount = 0;  // the own king's candidates (contiguous, <= 4) on regular boards
  if (!exotic) {
    // ---- regular position: bitmask legality (suicide_fast) for non-king moves
    FastCtx f;
    {
      const int es = -player;
      f.K = ownK; f.kr = ownK / 9; f.kc = ownK - f.kr * 9; f.geo = player;
      f.touch = g_touch + ownK * 90;
      f.rowm = w.rows[f.kr];
      f.colm = w.cols[f.kc];
      f.er_row = f.ec_row = f.ep_row = f.er_col = f.ec_col = f.ep_col = 0;
      // 9 row squares then 10 column squares of the king, L at a time
#pragma unroll 1
      for (int base = 0; base < 19; base += L) {
        const int t = base + lane;
        int q = 0;
        if (t < 9) q = w.sq[f.kr * 9 + t];
        else if (t < 19) q = w.sq[(t - 9) * 9 + f.kc];
        const uint64_t br = (uint64_t)T::ballot(q == es * ROOK) << base;
        const uint64_t bc = (uint64_t)T::ballot(q == es * CANNON) << base;
        const uint64_t bp = (uint64_t)T::ballot(q == es * PAWN) << base;
        f.er_row |= (unsigned)br & 0x1FFu; f.ec_row |= (unsigned)bc & 0x1FFu; f.ep_row |= (unsigned)bp & 0x1FFu;
        f.er_col |= (unsigned)(br >> 9) & 0x3FFu; f.ec_col |= (unsigned)(bc >> 9) & 0x3FFu;
        f.ep_col |= (unsigned)(bp >> 9) & 0x3FFu;
      }
      // tasks 0..7: knight squares; 8..11: legs
      f.ekn = f.legocc = 0;
#pragma unroll 1
      for (int base = 0; base < 12; base += L) {
        const int t = base + lane;
        bool flag = false;
        if (t < 12) {
          const int d = t & 3;  // tasks 0-3: K+(2a,b), 4-7: K+(a,2b), 8-11: the leg K+(a,b)
          const int a = (d & 2) ? 1 : -1, b = (d & 1) ? 1 : -1;
          const int lr = f.kr + a, lc = f.kc + b;
          const bool lon = lr >= 0 && lr <= 9 && lc >= 0 && lc <= 8;
          if (t >= 8) {
            flag = !lon || w.sq[lr * 9 + lc] != 0;
          } else if (lon) {
            const int r2 = (t & 4) ? lr : f.kr + 2 * a, c2 = (t & 4) ? f.kc + 2 * b : lc;
            flag = r2 >= 0 && r2 <= 9 && c2 >= 0 && c2 <= 8 && w.sq[r2 * 9 + c2] == es * KNIGHT;
          }
        }
        const unsigned b2 = T::ballot(flag) << base;
        f.ekn |= b2 & 0xFFu;
        f.legocc |= (b2 >> 8) & 0xFu;
      }
      finish_fast_ctx(f, player, player == 1 ? g.black_king : g.red_king);
    }
    // B.1: classify — king move / touches the king's lines / cannot matter
    int nwl = 0;
    bool any_irrelevant = false;
#pragma unroll 1
    for (int base = 0; base < ncand; base += L) {
      const int j = base + lane;
      bool rel = false, isk = false;
      const bool valid = j < ncand;
      if (valid) {
        const int c = w.cand[j], from = c >> 8, to = c & 0x7f;
        isk = from == ownK;
        rel = !isk && (touches(f.kr, f.kc, from) || touches(f.kr, f.kc, to));
        if (!rel && !isk) w.cand[j] = (uint16_t)(c | kCandIrrelevant);
      }
      const unsigned m = T::ballot(rel), mk = T::ballot(isk);
      if (rel) w.wl[nwl + __popc(m & lt)] = (uint16_t)j;
      nwl += __popc(m);
      if (mk && kcount == 0) kfirst = base + __ffs(mk) - 1;
      kcount += __popc(mk);
      any_irrelevant |= T::any(valid && !rel && !isk);
    }
    if (any_irrelevant && lane == 0) w.wl[nwl] = kWlSentinel;
    nwl += any_irrelevant ? 1 : 0;
    T::sync();
    // B.2a: bitmask test of the worklist (the sentinel evaluates the position itself)
#pragma unroll 1
    for (int base = 0; base < nwl; base += L) {
      const int i = base + lane;
      bool is_cur = false, bad = false;
      if (i < nwl) {
        const int item = w.wl[i];
        is_cur = item == kWlSentinel;
        const int c = is_cur ? 0 : (int)w.cand[item];
        bad = suicide_fast(f, is_cur ? -1 : (c >> 8), is_cur ? -1 : (c & 0x7f));
        if (bad && !is_cur) w.cand[item] = (uint16_t)(c | kCandIllegal);
      }
      cur_bad |= T::any(is_cur && bad);
    }
  } else {
    legality_generic(&w, &g, ncand, lane, L);
  }
  // Probe round — ONE inlined copy of the 8-probe attack test per kernel.  Item 0: is this
  // side's king attacked under the PREVIOUS mover's geometry, i.e. make_move's is_checking
  // (:317; evaluated here because it needs the same post-move board).  Items 1..kcount: the own
  // king's moves on regular boards (8 probes at the new square + kings facing, :448-451).
  {
    const bool want_check = checked_out != nullptr && ownK >= 0;
    const int n_it = 1 + kcount;
    bool checked = false;
#pragma unroll 1
    for (int base = 0; base < n_it * 8; base += L) {
      const int idx = base + lane, it = idx >> 3, pr = idx & 7;
      bool active = it < n_it, hit = false;
      int K = ownK, geo = -player, from = -1, to = -1, mover = 0;
      if (it == 0) {
        active = want_check;
      } else if (active) {
        to = w.cand[kfirst + it - 1] & 0x7f;
        K = to; geo = player; from = ownK; mover = player * KING;
      }
      if (active) {
        const Probe p = make_probe(w, K, -player, geo, from, to, mover);
        hit = pr < 4 ? probe_ray(w, p, pr) : probe_diag(w, p, pr - 4, it == 0);
        if (it > 0 && pr == 0) {  // kings facing after the king's own move
          const int ek = player == 1 ? g.black_king : g.red_king;
          if (ek >= 0 && ek % 9 == p.kc) {
            const int er = ek / 9, lo = min(er, p.kr), hi = max(er, p.kr);
            hit |= (p.colm & (((1u << hi) - 1u) & ~((2u << lo) - 1u))) == 0;
          }
        }
      }
      const unsigned bal = T::ballot(active && hit);
      if (base == 0) checked = (bal & 0xFFu) != 0;
      if (active && it > 0 && pr == 0 && ((bal >> (lane & ~7)) & 0xFFu))
        w.cand[kfirst + it - 1] |= kCandIllegal;
    }
    if (checked_out) *checked_out = checked;
  }
  T::sync();

  // Phase B.3: ordered compaction
  int n_legal = 0;
#pragma unroll 1
  for (int base = 0; base < ncand; base += L) {
    const int j = base + lane;
    bool ok = false;
    int c = 0;
    if (j < ncand) {
      c = w.cand[j];
      ok = (c & kCandIrrelevant) ? !cur_bad : !(c & kCandIllegal);
    }
    const unsigned m = T::ballot(ok);
    const int idx = n_legal + __popc(m & lt);
    if (ok && idx < XQ_MAX_MOVES) w.moves[idx] = (int16_t)(((c >> 8) & 0x7f) * 90 + (c & 0x7f));
    n_legal += __popc(m);
  }
  if (n_legal > XQ_MAX_MOVES) {
    n_legal = XQ_MAX_MOVES;
    g.flags |= XQ_F_OVERFLOW;
  }
  T::sync();
  return n_legal;
}
#endif  // __CUDACC__

// ---- step -------------------------------------------------------------------
// _evaluate_position_change (chess_env.py:683-737); before the side switch.
XQ_HD double position_change(int type, int player, int from, int to,
                                                  int enemy_king) {
  const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
  double score = 0.0;
  const int advance = player == 1 ? fr - tr : tr - fr;
  if (advance > 0) {
    if (type == PAWN) score = xq_dadd(score, xq_dmul((double)advance, 2.0));
    else if (type == ROOK || type == CANNON) score = xq_dadd(score, xq_dmul((double)advance, 1.5));
    else if (type == KNIGHT) score = xq_dadd(score, (double)advance);
  }
  if (tc >= 3 && tc <= 5) {
    score = xq_dadd(score, 1.5);
    if (tr >= 3 && tr <= 6) score = xq_dadd(score, 1.0);
  }
  if (type == PAWN && (player == 1 ? tr < 5 : tr >= 5)) score = xq_dadd(score, 3.0);
  if (enemy_king >= 0) {
    const int kr = enemy_king / 9, kc = enemy_king - kr * 9;
    const int od = xq_abs(fr - kr) + xq_abs(fc - kc), nd = xq_abs(tr - kr) + xq_abs(tc - kc);
    if (nd < od) score = xq_dadd(score, xq_dmul((double)(od - nd), 0.5));
  }
  return score;
}

#if defined(__CUDACC__)
struct StepOut {
  double reward;  // valid on every lane (warp-uniform inputs)
  int is_int;
  int done;
  int n_next;     // legal moves of the new side to move (in w.moves) or -1 if not generated
  uint64_t key_next;  // position key of the new board ‖ new side to move
  int from, to, moving, captured;  // the applied move (for the deferred positional reward)
};

// make_move part 1 (chess_env.py:253-314,:338,:348-349): apply, caches, capture reward,
// position-history append, side switch.  o.done is set only by a king capture (:292-297).
// The check test (:317) needs the post-move board too and is evaluated by the movegen that
// follows (one probe round serves both); its consequences (:318-345) are in step_finish.
template <int L>
__device__ __forceinline__ StepOut step_apply(WarpSmem& w, Game& g, int move,
                                              uint64_t* __restrict__ hist, int hist_cap) {
  using T = Tile<L>;
  const int lane = T::lane();
  const int from = move / 90, to = move - from * 90;
  const int captured = w.sq[to], moving = w.sq[from];  // :265-266
  T::sync();
  if (lane == 0) {
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
  {  // incremental position key: three (piece, square) keys on three lanes, one mix64 site
    const int kp = lane == 2 ? captured : moving;
    const int ks = lane == 0 ? from : to;
    uint64_t x = (lane < 3 && kp != 0) ? piece_key(kp, ks) : 0ULL;
    x ^= T::shfl_xor(x, 1);
    x ^= T::shfl_xor(x, 2);
    g.bkey ^= T::shfl(x, 0);
  }
  T::sync();

  if (moving == KING) g.red_king = to;  // :271-279
  else if (moving == -KING) g.black_king = to;
  if (captured == KING) g.red_king = -1;
  else if (captured == -KING) g.black_king = -1;
  g.no_capture = captured != 0 ? 0 : g.no_capture + 1;  // :282-285

  StepOut o;
  o.from = from; o.to = to; o.moving = moving; o.captured = captured;
  o.reward = 0.0;
  o.is_int = 1;
  o.done = 0;
  const int acap = captured < 0 ? -captured : captured;
  if (acap == KING) {  // :292-297
    g.winner = g.player;
    o.reward = 100.0;
    o.done = 1;
    g.reason = XQ_REASON_KING_CAPTURE;
    g.done = 1;
  } else if (captured != 0) {  // :300-314
    const double base = acap == ROOK ? 9.0 : acap == CANNON ? 4.5 : acap == KNIGHT ? 4.0
                        : (acap == BISHOP || acap == ADVISOR) ? 2.0 : acap == PAWN ? 1.0 : 0.0;
    o.reward = xq_dmul(base, 2.0);
    o.is_int = 0;
    if (acap == ADVISOR || acap == BISHOP) o.reward = xq_dadd(o.reward, 3.0);
  }
  if (g.hist_len < hist_cap) {  // :338, stored with the MOVER's side byte
    if (lane == 0) hist[g.hist_len] = g.bkey ^ side_key(g.player);
    g.hist_len += 1;
  } else {
    g.flags |= XQ_F_OVERFLOW;
  }
  g.player = -g.player;  // :348-349
  g.move_count += 1;
  o.key_next = g.bkey ^ side_key(g.player);
  o.n_next = -1;
  T::sync();
  return o;
}

// make_move part 2.  `checking` = is_checking of :317 (false after a king capture, whose cache
// is None): check bonus / consecutive_checks / positional reward (:318-335), check_history
// append (:341); then, unless the king was captured, the terminal chain for the side now to
// move given its legal-move count (:352-397) and the 70-ply cap (:400-404).
template <int L>
__device__ __forceinline__ void step_finish(const WarpSmem& w, Game& g, StepOut& o, int n_legal,
                                            bool checking, const uint64_t* __restrict__ hist) {
  using T = Tile<L>;
  const int lane = T::lane();
  const int mover = -g.player;
  if (!o.done && checking) {  // :318-327
    if (g.cchecks == 0) { o.reward = xq_dadd(o.reward, 15.0); o.is_int = 0; }
    else if (g.cchecks == 1) { o.reward = xq_dadd(o.reward, 10.0); o.is_int = 0; }
    else if (g.cchecks == 2) { o.reward = xq_dadd(o.reward, 5.0); o.is_int = 0; }
    g.cchecks += 1;
  } else {  // :328-335
    g.cchecks = 0;
    if (o.captured == 0 && !o.done) {
      const int ek = mover == 1 ? g.black_king : g.red_king;
      const double pcg = position_change(o.moving < 0 ? -o.moving : o.moving, mover, o.from, o.to, ek);
      o.reward = xq_dadd(o.reward, xq_dmul(pcg, 0.01));
      o.is_int = 0;
    }
  }
  g.check_bits = (g.check_bits << 1) | (checking ? 1u : 0u);  // :341
  g.check_len += 1;
  if (o.done) return;  // king capture: no terminal chain (:352), no move cap (:400 "not done")

  o.n_next = n_legal;
  const bool chk_now = n_legal == 0 ? in_check_cold(&w, &g, g.player) : false;
  if (n_legal == 0 && chk_now) {  // :354, :614-628
    o.done = 1; o.reward = 200.0; o.is_int = 1;
    g.winner = -g.player;
    g.reason = XQ_REASON_CHECKMATE;
  } else {
    int cnt = 0;  // :362, :598-605 — query uses the NEW side byte (quirk A.7)
#pragma unroll 1
    for (int i = lane; i < g.hist_len; i += L) cnt += hist[i] == o.key_next;
    cnt = (int)T::sum((unsigned)cnt);
    if (cnt >= 3) {
      o.done = 1; o.reward = 0.0; o.is_int = 1;
      g.winner = 0;
      g.reason = XQ_REASON_REPETITION;
    } else if (g.no_capture >= 100) {  // :369, :612
      o.done = 1; o.reward = 0.0; o.is_int = 1;
      g.winner = 0;
      g.reason = XQ_REASON_FIFTY;
    } else if (n_legal == 0) {  // :376, :630-644
      o.done = 1; o.reward = 100.0; o.is_int = 1;
      g.winner = -g.player;
      g.reason = XQ_REASON_STALEMATE;
    } else if (g.check_len >= 12 && __popc(g.check_bits & 0xFFFu) >= 10) {  // :384, :646-662
      o.done = 1; o.reward = -10.0; o.is_int = 1;
      g.winner = -g.player;
      g.reason = XQ_REASON_PERPETUAL_CHECK;
    }  // :392 perpetual chase never fires (:674)
  }
  if (!o.done && g.move_count >= 70) {  // :400-404
    o.done = 1; o.reward = -2.0; o.is_int = 1;
    g.winner = 0;
    g.reason = XQ_REASON_MOVE_CAP;
  }
  if (o.done) g.done = 1;
}

// make_move (chess_env.py:253-406).  hist: this game's position_history row.
template <int L>
__device__ __forceinline__ StepOut step(WarpSmem& w, Game& g, int move, uint64_t* __restrict__ hist,
                                        int hist_cap, const uint32_t* __restrict__ leap) {
  StepOut o = step_apply<L>(w, g, move, hist, hist_cap);
  bool checking = false;
  int n_legal = -1;
  if (!o.done) n_legal = movegen<L>(w, g, leap, &checking);  // :317 + :354/:376 in one pass
  step_finish<L>(w, g, o, n_legal, checking, hist);
  return o;
}

__device__ __forceinline__ uint8_t step_flags(const Game& g, const StepOut& o) {
  return (uint8_t)((o.done & 1) | ((o.is_int & 1) << 1) | (((g.winner + 1) & 3) << 2) |
                   ((g.reason & 15) << 4));
}

// ---- philox4x32-10 pick ------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Capture-biased pick (diagnostic workloads only): kept out of line so that the uniform-pick
// hot loop does not carry its code.  lane/stride/member mask describe the calling tile.
static __device__ __noinline__ int pick_capture(const WarpSmem* wp, int n, uint32_t x0, int lane,
                                                int stride, unsigned mask, int shift) {
  const WarpSmem& w = *wp;
  const unsigned lanes = stride == 32 ? 0xffffffffu : ((1u << stride) - 1u);
  int ncap = 0;
  for (int base = 0; base < n; base += stride) {
    const int i = base + lane;
    const bool cap = i < n && w.sq[(int)w.moves[i] % 90] != 0;
    ncap += __popc((__ballot_sync(mask, cap) >> shift) & lanes);
  }
  if (ncap == 0) return -1;
  int k = (int)(x0 % (uint32_t)ncap);
  for (int base = 0; base < n; base += stride) {
    const int i = base + lane;
    const bool cap = i < n && w.sq[(int)w.moves[i] % 90] != 0;
    const unsigned m = (__ballot_sync(mask, cap) >> shift) & lanes;
    const int c = __popc(m);
    if (k < c) return base + (int)__fns(m, 0, k + 1);
    k -= c;
  }
  return -1;
}

// Index into w.moves[0..n) chosen by the shared pick rule.
template <int L>
__device__ __forceinline__ int pick_index(const WarpSmem& w, int n, uint64_t seed, uint32_t game_id,
                                          uint32_t ply, int capture_bias) {
  uint32_t x[4];
  philox4x32(game_id, ply, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  if (capture_bias > 0 && (int)(x[1] & 0xFFu) < capture_bias) {
    const int idx = pick_capture(&w, n, x[0], Tile<L>::lane(), L, Tile<L>::mask(), Tile<L>::shift());
    if (idx >= 0) return idx;
  }
  return (int)(x[0] % (uint32_t)n);
}

#endif  // __CUDACC__

}  // namespace xq
