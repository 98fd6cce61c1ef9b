// xq_mcts.cu — MCTS select / expand / backup over a flat per-game node pool
// (self_play.py:19-175), one warp per tree, sm_100a.
//
// Layout of one tree (xq_mcts_tree_bytes(n) bytes, 16-byte aligned):
//   TreeHeader                       64 B
//   Node nodes[1 + 128*waves]        32 B each; children of a node contiguous,
//                                    in legal-move order (MCTSNode.children dict order)
//   StateRec states[waves + 2]       board 96 B + xq_meta 32 B + path history u64[waves+1]
// waves = ceil(n/8).  At most one node is expanded per wave, so
// the pools cannot overflow.  Expanded nodes keep their env state: a simulation
// walks the tree with PUCT only and steps the rules engine once, at its leaf.
// No atomics: a tree is touched by exactly one warp.
#include <cuda_bf16.h>

#include "xq_rules.cuh"

namespace xq {

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

constexpr int kWave = 8;  // self_play.py:101
constexpr int kThreadsM = kWarpsPerCta * 32;

struct TreeHeader {
  int32_t sims_total, sims_done, n_nodes, n_states;
  int32_t pending_node, pending_m, node_cap, state_cap;
  int32_t hist_cap, overflow, pad[6];
};
static_assert(sizeof(TreeHeader) == 64, "TreeHeader");

struct __align__(16) Node {
  double value_sum;     // self_play.py:27
  int32_t visits;       // :26
  float prior;          // :28 (numpy.float32)
  int32_t first_child;  // -1 = leaf (:36-38)
  int32_t parent;       // :22
  int16_t n_children;
  int16_t move;         // :23
  int16_t state;        // index into states[] once the env state is known
  int8_t term;          // 0 unknown, 1 non-terminal, 2 terminal leaf
  int8_t tval;          // terminal value +1/-1/0 (:128-133)
};
static_assert(sizeof(Node) == 32, "Node");

struct TreeDims {
  int waves, node_cap, state_cap, hist_cap;
  int64_t node_off, state_off, state_bytes, bytes;
};

__host__ __device__ inline TreeDims tree_dims(int n_sims) {
  TreeDims d;
  d.waves = (n_sims + kWave - 1) / kWave;
  if (d.waves < 1) d.waves = 1;
  d.node_cap = 1 + XQ_MAX_MOVES * d.waves;
  d.state_cap = d.waves + 2;
  d.hist_cap = d.waves + 1;
  d.node_off = sizeof(TreeHeader);
  d.state_off = d.node_off + (int64_t)d.node_cap * sizeof(Node);
  d.state_bytes = XQ_BOARD_STRIDE + sizeof(xq_meta) + (int64_t)d.hist_cap * 8;
  d.state_bytes = (d.state_bytes + 15) & ~15LL;
  d.bytes = d.state_off + d.state_bytes * d.state_cap;
  d.bytes = (d.bytes + 63) & ~63LL;
  return d;
}

struct Tree {
  TreeHeader* h;
  Node* nodes;
  char* states;
  int64_t state_bytes;
  __device__ int8_t* board(int s) const { return reinterpret_cast<int8_t*>(states + s * state_bytes); }
  __device__ xq_meta* meta(int s) const {
    return reinterpret_cast<xq_meta*>(states + s * state_bytes + XQ_BOARD_STRIDE);
  }
  __device__ uint64_t* hist(int s) const {
    return reinterpret_cast<uint64_t*>(states + s * state_bytes + XQ_BOARD_STRIDE + sizeof(xq_meta));
  }
};

__device__ __forceinline__ Tree tree_at(void* trees, const TreeDims& d, int g) {
  char* base = reinterpret_cast<char*>(trees) + (int64_t)g * d.bytes;
  Tree t;
  t.h = reinterpret_cast<TreeHeader*>(base);
  t.nodes = reinterpret_cast<Node*>(base + d.node_off);
  t.states = base + d.state_off;
  t.state_bytes = d.state_bytes;
  return t;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsM)
    mcts_init_kernel(void* trees, int n_sims, const int8_t* __restrict__ board,
                     const xq_meta* __restrict__ meta, const uint8_t* __restrict__ active,
                     int n_games) {
  const int g = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (g >= n_games) return;
  const int lane = threadIdx.x & 31;
  const TreeDims d = tree_dims(n_sims);
  Tree t = tree_at(trees, d, g);
  const bool on = active ? active[g] != 0 : true;
  if (lane < XQ_BOARD_STRIDE / 4)
    reinterpret_cast<uint32_t*>(t.board(0))[lane] =
        reinterpret_cast<const uint32_t*>(board + (size_t)g * XQ_BOARD_STRIDE)[lane];
  if (lane == 0) {
    TreeHeader h = {};
    h.sims_total = on ? n_sims : 0;
    h.n_nodes = 1;
    h.n_states = 1;
    h.pending_node = -1;
    h.node_cap = d.node_cap;
    h.state_cap = d.state_cap;
    h.hist_cap = d.hist_cap;
    *t.h = h;
    Node r = {};
    r.first_child = -1;
    r.parent = -1;
    r.move = -1;
    r.state = 0;
    t.nodes[0] = r;
    // MCTS._copy_env (self_play.py:156-175)
    xq_meta m = meta[g];
    xq_meta c = {};
    c.player = m.player;
    c.winner = m.winner;
    c.red_king = m.red_king;
    c.black_king = m.black_king;
    c.move_count = m.move_count;
    c.no_capture = m.no_capture;
    *t.meta(0) = c;
  }
}

// select_child (self_play.py:40-59): float32 PUCT, strict '>' => first max wins.
__device__ __forceinline__ int puct_select(const Tree& t, int node) {
  const int lane = threadIdx.x & 31;
  const Node nd = t.nodes[node];
  const float sq = (float)sqrt((double)nd.visits);
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = lane; i < nd.n_children; i += 32) {
    const Node* c = &t.nodes[nd.first_child + i];
    const uint4 raw = *reinterpret_cast<const uint4*>(c);  // value_sum, visits, prior
    const double vs = __hiloint2double((int)raw.y, (int)raw.x);
    const int n = (int)raw.z;
    const float p = __uint_as_float(raw.w);
    const float q = n == 0 ? 0.0f : __double2float_rn(vs / (double)n);  // :30-34
    float u = __fmul_rn(1.5f, p);
    u = __fmul_rn(u, sq);
    u = __fdiv_rn(u, (float)(1 + n));
    const float s = __fadd_rn(q, u);
    if (s > best) {
      best = s;
      best_i = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(kFull, best, o);
    const int oi = __shfl_xor_sync(kFull, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  return nd.first_child + best_i;
}

// MCTSNode.update (self_play.py:70-80): sequential float64 adds up the parent chain.
__device__ __forceinline__ void backup_path(const Tree& t, int node, double v) {
  while (node >= 0) {
    Node* nd = &t.nodes[node];
    nd->visits += 1;
    nd->value_sum = __dadd_rn(nd->value_sum, v);
    v = -v;
    node = nd->parent;
  }
}

__global__ void __launch_bounds__(kThreadsM, 4)
    mcts_select_kernel(void* trees, int n_sims, int wave_size, int8_t* __restrict__ leaf_board,
                       int8_t* __restrict__ leaf_player, int16_t* __restrict__ leaf_moves,
                       int16_t* __restrict__ leaf_n, int16_t* __restrict__ leaf_mult, int n_games) {
  __shared__ WarpSmem slab[kWarpsPerCta];
  const int g = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (g >= n_games) return;
  WarpSmem& w = slab[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const TreeDims d = tree_dims(n_sims);
  Tree t = tree_at(trees, d, g);
  TreeHeader h = *t.h;
  const int count = min(wave_size, h.sims_total - h.sims_done);  // :104-105
  int pending = -1, mult = 0, pend_n = 0;
  for (int k = 0; k < count; ++k) {
    int node = 0;
    while (true) {  // :117-119
      const Node nd = t.nodes[node];
      if (nd.first_child >= 0) {
        node = puct_select(t, node);
        continue;
      }
      int term = nd.term, tval = nd.tval;
      if (term == 0) {  // first visit: derive this node's env state
        Game G;
        int n_legal;
        int slot;
        if (nd.parent < 0) {  // root: state 0 was written by init
          slot = 0;
          load_board<32>(w, t.board(0));
          build_masks<32>(w);
          G = load_meta(t.meta(0));
          n_legal = movegen<32>(w, G, g_leap);  // :123
        } else {
          const int ps = t.nodes[nd.parent].state;
          slot = h.n_states;  // scratch until proven non-terminal
          load_board<32>(w, t.board(ps));
          build_masks<32>(w);
          G = load_meta(t.meta(ps));
          G.bkey = board_key<32>(w);
          uint64_t* hs = t.hist(slot);
          const uint64_t* hp = t.hist(ps);
          for (int i = lane; i < G.hist_len; i += 32) hs[i] = hp[i];
          __syncwarp();
          const StepOut o = step<32>(w, G, nd.move, hs, d.hist_cap, g_leap);  // :119
          n_legal = o.n_next < 0 ? 0 : o.n_next;
          if (o.n_next < 0 && G.winner == XQ_WINNER_NONE) n_legal = 0;
        }
        if (n_legal == 0 || G.winner != XQ_WINNER_NONE) {  // :126-133
          term = 2;
          tval = G.winner == G.player ? 1 : (G.winner == -G.player ? -1 : 0);
        } else {
          term = 1;
          store_board<32>(w, t.board(slot));
          store_meta<32>(t.meta(slot), G);
          if (slot != 0) h.n_states += 1;
          int16_t* lm = leaf_moves + (size_t)g * XQ_MAX_MOVES;
          for (int i = lane; i < n_legal; i += 32) lm[i] = w.moves[i];
          pend_n = n_legal;
          if (lane == 0) t.nodes[node].state = (int16_t)slot;
        }
        if (lane == 0) {
          t.nodes[node].term = (int8_t)term;
          t.nodes[node].tval = (int8_t)tval;
        }
        __syncwarp();
      }
      if (term == 2) {  // :134-135
        if (lane == 0) backup_path(t, node, (double)tval);
        __syncwarp();
      } else {  // :138-139 — every remaining sim of the wave reaches this same leaf
        pending = node;
        mult = count - k;
      }
      break;
    }
    if (pending >= 0) break;
  }
  // leaf outputs
  if (pending >= 0) {
    const int s = t.nodes[pending].state;
    if (pend_n == 0) {  // unreachable: a non-terminal leaf is materialised in the wave that finds it
      h.overflow |= 2;
    }
    if (lane < XQ_BOARD_STRIDE / 4)
      reinterpret_cast<uint32_t*>(leaf_board + (size_t)g * XQ_BOARD_STRIDE)[lane] =
          reinterpret_cast<const uint32_t*>(t.board(s))[lane];
    if (lane == 0) leaf_player[g] = t.meta(s)->player;
  }
  if (lane == 0) {
    leaf_n[g] = (int16_t)(pending >= 0 ? pend_n : 0);
    leaf_mult[g] = (int16_t)mult;
    h.sims_done += count > 0 ? count : 0;
    h.pending_node = pending;
    h.pending_m = mult;
    *t.h = h;
  }
}

// row_of_game (optional): priors / values are COMPACT arrays that hold only the games whose wave
// reached a network leaf; row_of_game[g] is game g's row in them (xq_compact_leaves).
template <typename V>
__global__ void __launch_bounds__(kThreadsM)
    mcts_backup_kernel(void* trees, int n_sims, const int16_t* __restrict__ leaf_moves,
                       const int16_t* __restrict__ leaf_n, const float* __restrict__ priors,
                       const V* __restrict__ values, int values_per_game,
                       const int32_t* __restrict__ row_of_game, int n_games) {
  const int g = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (g >= n_games) return;
  const int lane = threadIdx.x & 31;
  const TreeDims d = tree_dims(n_sims);
  Tree t = tree_at(trees, d, g);
  TreeHeader h = *t.h;
  const int node = h.pending_node;
  if (node < 0) return;
  const int n = leaf_n[g];
  const size_t row = row_of_game ? (size_t)row_of_game[g] : (size_t)g;
  if (t.nodes[node].first_child < 0) {  // expand (:61-68); idempotent
    const int first = h.n_nodes;
    if (first + n > h.node_cap) {
      if (lane == 0) {
        h.overflow |= 1;
        h.pending_node = -1;
        *t.h = h;
      }
      return;
    }
    for (int i = lane; i < n; i += 32) {
      Node c = {};
      c.prior = priors[row * XQ_MAX_MOVES + i];
      c.first_child = -1;
      c.parent = node;
      c.move = leaf_moves[(size_t)g * XQ_MAX_MOVES + i];
      c.state = -1;
      t.nodes[first + i] = c;
    }
    if (lane == 0) {
      t.nodes[node].first_child = first;
      t.nodes[node].n_children = (int16_t)n;
    }
    h.n_nodes = first + n;
  }
  __syncwarp();
  if (lane == 0) {  // one update per queued simulation, in queue order (:146-148)
    for (int k = 0; k < h.pending_m; ++k) {
      const double v = (double)values[values_per_game == 1 ? row : row * kWave + k];
      backup_path(t, node, v);
    }
    h.pending_node = -1;
    h.pending_m = 0;
    *t.h = h;
  }
}

__global__ void __launch_bounds__(kThreadsM)
    mcts_root_visits_kernel(const void* trees, int n_sims, int16_t* __restrict__ moves,
                            int32_t* __restrict__ visits, int16_t* __restrict__ n_children,
                            int n_games) {
  const int g = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (g >= n_games) return;
  const int lane = threadIdx.x & 31;
  const TreeDims d = tree_dims(n_sims);
  Tree t = tree_at(const_cast<void*>(trees), d, g);
  const Node root = t.nodes[0];
  const int n = root.first_child >= 0 ? root.n_children : 0;
  for (int i = lane; i < XQ_MAX_MOVES; i += 32) {
    const bool ok = i < n;
    moves[(size_t)g * XQ_MAX_MOVES + i] = ok ? t.nodes[root.first_child + i].move : (int16_t)-1;
    visits[(size_t)g * XQ_MAX_MOVES + i] = ok ? t.nodes[root.first_child + i].visits : 0;
  }
  if (lane == 0) n_children[g] = (int16_t)n;
}

// Temperature sampling of the move to play (self_play.py:219-243), one warp per game: the lanes
// load the visit counts coalesced and do the expensive part (pow, float64 division) in parallel;
// the two float64 sums run in index order, as numpy's sum/cumsum do, on values passed round by
// shuffle — every lane computes the same sums, so the early exit is warp-uniform.
__global__ void __launch_bounds__(256)
    sample_moves_kernel(const int32_t* __restrict__ visits, const int16_t* __restrict__ n_children,
                        const uint8_t* __restrict__ active, double temperature, uint64_t seed,
                        uint32_t first_game_id, uint32_t ply, int16_t* __restrict__ chosen,
                        int n_games) {
  const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= n_games) return;
  constexpr int kPer = XQ_MAX_MOVES / 32;
  const int n = min((int)n_children[g], XQ_MAX_MOVES);
  if (n <= 0 || (active && !active[g])) {
    if (lane == 0) chosen[g] = -1;
    return;
  }
  const int32_t* v = visits + (size_t)g * XQ_MAX_MOVES;
  int cnt[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) cnt[j] = 32 * j + lane < n ? v[32 * j + lane] : 0;
  uint32_t x[4];
  philox4x32(first_game_id + (uint32_t)g, ply, 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  int pick = 0;
  if (temperature < 0.01) {  // :224-227, np.argmax = first maximum
    int best = -1, best_i = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j)
      if (32 * j + lane < n && cnt[j] > best) { best = cnt[j]; best_i = 32 * j + lane; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    pick = best_i;
  } else {
    const double inv_t = 1.0 / temperature;
    // counts ** (1/T): exact shortcuts for the reference's two temperatures (trainer.py:166)
    double w[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const double c = (double)cnt[j];
      w[j] = inv_t == 1.0 ? c : (inv_t == 2.0 ? __dmul_rn(c, c) : pow(c, inv_t));
    }
    double total = 0.0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (32 * j >= n) break;
#pragma unroll 1
      for (int l = 0; l < 32 && 32 * j + l < n; ++l) total = __dadd_rn(total, __shfl_sync(0xffffffffu, w[j], l));
    }
    if (!(total > 0.0)) {  // all-zero counts (n_sims <= 8): the reference divides by zero here
      pick = (int)(x[0] % (uint32_t)n);
    } else {
      const double u = (double)((((uint64_t)x[0] << 32) | x[1]) >> 11) * (1.0 / 9007199254740992.0);
#pragma unroll
      for (int j = 0; j < kPer; ++j) w[j] = __ddiv_rn(w[j], total);
      double cdf = 0.0;
      pick = n - 1;
      bool found = false;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {  // searchsorted(cumsum(p), u, side="right")
        if (found || 32 * j >= n) break;
#pragma unroll 1
        for (int l = 0; l < 32 && 32 * j + l < n; ++l) {
          cdf = __dadd_rn(cdf, __shfl_sync(0xffffffffu, w[j], l));
          if (cdf > u) { pick = 32 * j + l; found = true; break; }
        }
      }
    }
  }
  if (lane == 0) chosen[g] = (int16_t)pick;
}

// Game-loop glue of one ply (see xq_selfplay_commit / xq_selfplay_finish in the header):
// one CTA of XQ_MAX_MOVES threads per game.
__global__ void __launch_bounds__(XQ_MAX_MOVES)
    selfplay_commit_kernel(const int16_t* __restrict__ root_moves, const int32_t* __restrict__ root_visits,
                           const int16_t* __restrict__ root_n, const int16_t* __restrict__ chosen,
                           const int8_t* __restrict__ board, const xq_meta* __restrict__ meta,
                           int8_t* __restrict__ rec_board, int8_t* __restrict__ rec_player,
                           int16_t* __restrict__ rec_moves, int32_t* __restrict__ rec_visits,
                           int16_t* __restrict__ rec_n, uint8_t* __restrict__ rec_played,
                           int16_t* __restrict__ rec_move, int16_t* __restrict__ move,
                           int32_t* __restrict__ any_active) {
  const int g = blockIdx.x, k = threadIdx.x;
  const size_t row = (size_t)g * XQ_MAX_MOVES;
  rec_moves[row + k] = root_moves[row + k];
  rec_visits[row + k] = root_visits[row + k];
  if (k < XQ_NSQ) rec_board[(size_t)g * XQ_NSQ + k] = board[(size_t)g * XQ_BOARD_STRIDE + k];
  if (k == 0) {
    const int c = chosen[g];
    const int16_t mv = c >= 0 ? root_moves[row + c] : (int16_t)-1;
    rec_n[g] = root_n[g];
    rec_player[g] = meta[g].player;
    rec_played[g] = c >= 0 ? 1 : 0;
    rec_move[g] = mv;
    move[g] = mv;
    if (g == 0) *any_active = 0;
  }
}

__global__ void __launch_bounds__(256)
    selfplay_finish_kernel(const int16_t* __restrict__ move, const uint8_t* __restrict__ step_flags,
                           uint8_t* __restrict__ active, int32_t* __restrict__ any_active, int n) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  if (!active[g]) return;
  const bool on = move[g] >= 0 && !(step_flags[g] & XQ_STEP_DONE);
  active[g] = on ? 1 : 0;
  // number of games still running (zeroed by selfplay_commit_kernel): the host reads it a few
  // plies late as "any game left?" and as the upper bound for leaf compaction
  const unsigned m = __ballot_sync(__activemask(), on);
  if (on && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(any_active, __popc(m));
}

// Leaf compaction (ragged batches): ordered list of the games whose wave ended on a network
// leaf (leaf_n > 0), and each game's row in that list.  One CTA; every thread owns a contiguous
// slice, a block-wide exclusive scan of the slice counts places them in game order.
__global__ void __launch_bounds__(1024)
    compact_leaves_kernel(const int16_t* __restrict__ leaf_n, int n_games, int32_t* __restrict__ idx,
                          int32_t* __restrict__ row_of_game, int32_t* __restrict__ count) {
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x, per = (n_games + 1023) / 1024;
  const int lo = min(tid * per, n_games), hi = min(lo + per, n_games);
  int mine = 0;
  for (int g = lo; g < hi; ++g) mine += leaf_n[g] > 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = warp_sums[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, w, o);
      if (tid >= o) w += v;
    }
    warp_sums[tid] = w;
  }
  __syncthreads();
  int pos = incl - mine + ((tid >> 5) ? warp_sums[(tid >> 5) - 1] : 0);
  for (int g = lo; g < hi; ++g) {
    if (leaf_n[g] > 0) {
      idx[pos] = g;
      row_of_game[g] = pos++;
    } else {
      row_of_game[g] = -1;
    }
  }
  if (tid == 1023) *count = warp_sums[31];
}

// Rows idx[0..count) of the leaf arrays copied into compact arrays of `rows` rows; the padding
// rows (count..rows) get an empty position (no legal moves -> all-zero priors, ignored values).
__global__ void __launch_bounds__(128)
    gather_leaves_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ count, int rows,
                         const int8_t* __restrict__ leaf_board, const int8_t* __restrict__ leaf_player,
                         const int16_t* __restrict__ leaf_moves, const int16_t* __restrict__ leaf_n,
                         int8_t* __restrict__ out_board, int8_t* __restrict__ out_player,
                         int16_t* __restrict__ out_moves, int16_t* __restrict__ out_n) {
  const int r = blockIdx.x, k = threadIdx.x;
  if (r >= rows) return;
  const bool live = r < *count;
  const int g = live ? idx[r] : 0;
  // 96 board bytes = 24 words, 128 moves = 64 words
  if (k < XQ_BOARD_STRIDE / 4)
    reinterpret_cast<uint32_t*>(out_board + (size_t)r * XQ_BOARD_STRIDE)[k] =
        live ? reinterpret_cast<const uint32_t*>(leaf_board + (size_t)g * XQ_BOARD_STRIDE)[k] : 0u;
  if (k >= 32 && k < 32 + XQ_MAX_MOVES / 2)
    reinterpret_cast<uint32_t*>(out_moves + (size_t)r * XQ_MAX_MOVES)[k - 32] =
        live ? reinterpret_cast<const uint32_t*>(leaf_moves + (size_t)g * XQ_MAX_MOVES)[k - 32] : 0u;
  if (k == 127) {
    out_player[r] = live ? leaf_player[g] : (int8_t)1;
    out_n[r] = live ? leaf_n[g] : (int16_t)0;
  }
}

// Mirror of the oracle's deterministic evaluator (order-independent arithmetic).
__global__ void __launch_bounds__(kThreadsM)
    hash_eval_kernel(const int8_t* __restrict__ board, int board_stride,
                     const int8_t* __restrict__ player, const int16_t* __restrict__ moves,
                     const int16_t* __restrict__ n_moves, int flat, float* __restrict__ priors,
                     double* __restrict__ values, int n) {
  __shared__ WarpSmem slab[kWarpsPerCta];
  const int g = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (g >= n) return;
  WarpSmem& w = slab[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  for (int s = lane; s < XQ_NSQ; s += 32) w.sq[s] = board[(size_t)g * board_stride + s];
  __syncwarp();
  const uint64_t hsh = board_key<32>(w) ^ side_key(player[g]);
  const int cnt = n_moves[g];
  double wts[XQ_MAX_MOVES / 32];
  double sum = 0.0;
#pragma unroll
  for (int k = 0; k < XQ_MAX_MOVES / 32; ++k) {
    const int i = k * 32 + lane;
    double x = 0.0;
    if (i < cnt) {
      const uint64_t u =
          flat ? 0ULL
               : mix64(hsh ^ ((uint64_t)(uint16_t)moves[(size_t)g * XQ_MAX_MOVES + i] * 0x9E3779B97F4A7C15ULL));
      x = (double)((u >> 40) + 1);
    }
    wts[k] = x;
    sum += x;  // integers < 2^24: exact in any order
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
#pragma unroll
  for (int k = 0; k < XQ_MAX_MOVES / 32; ++k) {
    const int i = k * 32 + lane;
    priors[(size_t)g * XQ_MAX_MOVES + i] = i < cnt ? __double2float_rn(wts[k] / sum) : 0.0f;
  }
  if (lane == 0) values[g] = (double)((hsh >> 11) & 0xFFFFFu) / 524288.0 - 1.0;
}

}  // namespace xq

using namespace xq;

extern "C" {

#define XQM_REQUIRE(cond, msg)                                     \
  do {                                                             \
    if (!(cond)) return fail(XQ_E_ARG, "%s: %s", __func__, msg);   \
  } while (0)

static inline int ctas_m(int n) { return (n + kWarpsPerCta - 1) / kWarpsPerCta; }

int64_t xq_mcts_tree_bytes(int num_simulations) { return tree_dims(num_simulations).bytes; }

int xq_mcts_init(void* trees, int num_simulations, const int8_t* board, const xq_meta* meta,
                 const uint8_t* active, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(trees && board && meta && n_games > 0 && num_simulations > 0,
              "null pointer or non-positive size");
  mcts_init_kernel<<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
      trees, num_simulations, board, meta, active, n_games);
  return check_launch("xq_mcts_init");
}

int xq_mcts_select(void* trees, int num_simulations, int wave_size, int8_t* leaf_board,
                   int8_t* leaf_player, int16_t* leaf_moves, int16_t* leaf_n, int16_t* leaf_mult,
                   int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(trees && leaf_board && leaf_player && leaf_moves && leaf_n && leaf_mult && n_games > 0,
              "null pointer or non-positive size");
  XQM_REQUIRE(wave_size >= 1 && wave_size <= kWave && num_simulations > 0, "wave_size not in [1,8]");
  mcts_select_kernel<<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
      trees, num_simulations, wave_size, leaf_board, leaf_player, leaf_moves, leaf_n, leaf_mult,
      n_games);
  return check_launch("xq_mcts_select");
}

int xq_mcts_backup(void* trees, int num_simulations, const int16_t* leaf_moves,
                   const int16_t* leaf_n, const float* priors, const void* values, int values_f32,
                   int values_per_game, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(trees && leaf_moves && leaf_n && priors && values && n_games > 0,
              "null pointer or non-positive size");
  XQM_REQUIRE(values_per_game == 1 || values_per_game == kWave, "values_per_game must be 1 or 8");
  if (values_f32)
    mcts_backup_kernel<float><<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
        trees, num_simulations, leaf_moves, leaf_n, priors, (const float*)values, values_per_game,
        nullptr, n_games);
  else
    mcts_backup_kernel<double><<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
        trees, num_simulations, leaf_moves, leaf_n, priors, (const double*)values, values_per_game,
        nullptr, n_games);
  return check_launch("xq_mcts_backup");
}

int xq_mcts_backup_rows(void* trees, int num_simulations, const int16_t* leaf_moves,
                        const int16_t* leaf_n, const float* priors, const void* values, int values_f32,
                        const int32_t* row_of_game, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(trees && leaf_moves && leaf_n && priors && values && row_of_game && n_games > 0,
              "null pointer or non-positive size");
  if (values_f32)
    mcts_backup_kernel<float><<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
        trees, num_simulations, leaf_moves, leaf_n, priors, (const float*)values, 1, row_of_game, n_games);
  else
    mcts_backup_kernel<double><<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
        trees, num_simulations, leaf_moves, leaf_n, priors, (const double*)values, 1, row_of_game, n_games);
  return check_launch("xq_mcts_backup_rows");
}

int xq_compact_leaves(const int16_t* leaf_n, int n_games, int32_t* idx, int32_t* row_of_game,
                      int32_t* count, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(leaf_n && idx && row_of_game && count && n_games > 0, "null pointer or non-positive size");
  compact_leaves_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(leaf_n, n_games, idx, row_of_game, count);
  return check_launch("xq_compact_leaves");
}

int xq_gather_leaves(const int32_t* idx, const int32_t* count, int rows, const int8_t* leaf_board,
                     const int8_t* leaf_player, const int16_t* leaf_moves, const int16_t* leaf_n,
                     int8_t* out_board, int8_t* out_player, int16_t* out_moves, int16_t* out_n,
                     void* stream) {
  if (rows == 0) return 0;
  XQM_REQUIRE(idx && count && leaf_board && leaf_player && leaf_moves && leaf_n && out_board &&
                  out_player && out_moves && out_n && rows > 0,
              "null pointer or non-positive size");
  gather_leaves_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(idx, count, rows, leaf_board, leaf_player,
                                                              leaf_moves, leaf_n, out_board, out_player,
                                                              out_moves, out_n);
  return check_launch("xq_gather_leaves");
}

int xq_mcts_root_visits(const void* trees, int num_simulations, int16_t* moves, int32_t* visits,
                        int16_t* n_children, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(trees && moves && visits && n_children && n_games > 0 && num_simulations > 0,
              "null pointer or non-positive size");
  mcts_root_visits_kernel<<<ctas_m(n_games), kThreadsM, 0, (cudaStream_t)stream>>>(
      trees, num_simulations, moves, visits, n_children, n_games);
  return check_launch("xq_mcts_root_visits");
}

int xq_sample_moves(const int32_t* visits, const int16_t* n_children, const uint8_t* active,
                    double temperature, uint64_t seed, uint32_t first_game_id, uint32_t ply,
                    int16_t* chosen, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(visits && n_children && chosen && n_games > 0, "null pointer or non-positive size");
  XQM_REQUIRE(temperature >= 0.0, "negative temperature");
  sample_moves_kernel<<<(n_games + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
      visits, n_children, active, temperature, seed, first_game_id, ply, chosen, n_games);
  return check_launch("xq_sample_moves");
}

int xq_selfplay_commit(const int16_t* root_moves, const int32_t* root_visits, const int16_t* root_n,
                       const int16_t* chosen, const int8_t* board, const xq_meta* meta,
                       int8_t* rec_board, int8_t* rec_player, int16_t* rec_moves, int32_t* rec_visits,
                       int16_t* rec_n, uint8_t* rec_played, int16_t* rec_move, int16_t* move,
                       int32_t* any_active, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(root_moves && root_visits && root_n && chosen && board && meta && rec_board &&
                  rec_player && rec_moves && rec_visits && rec_n && rec_played && rec_move && move &&
                  any_active && n_games > 0,
              "null pointer or non-positive size");
  selfplay_commit_kernel<<<n_games, XQ_MAX_MOVES, 0, (cudaStream_t)stream>>>(
      root_moves, root_visits, root_n, chosen, board, meta, rec_board, rec_player, rec_moves,
      rec_visits, rec_n, rec_played, rec_move, move, any_active);
  return check_launch("xq_selfplay_commit");
}

int xq_selfplay_finish(const int16_t* move, const uint8_t* step_flags, uint8_t* active,
                       int32_t* any_active, int n_games, void* stream) {
  if (n_games == 0) return 0;
  XQM_REQUIRE(move && step_flags && active && any_active && n_games > 0,
              "null pointer or non-positive size");
  selfplay_finish_kernel<<<(n_games + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      move, step_flags, active, any_active, n_games);
  return check_launch("xq_selfplay_finish");
}

int xq_hash_eval(const int8_t* board, int board_stride, const int8_t* player, const int16_t* moves,
                 const int16_t* n_moves, int flat, float* priors, double* values, int n,
                 void* stream) {
  if (n == 0) return 0;
  XQM_REQUIRE(board && player && moves && n_moves && priors && values && n > 0 &&
                  board_stride >= XQ_NSQ,
              "null pointer or bad stride");
  hash_eval_kernel<<<ctas_m(n), kThreadsM, 0, (cudaStream_t)stream>>>(
      board, board_stride, player, moves, n_moves, flat, priors, values, n);
  return check_launch("xq_hash_eval");
}

}  // extern "C"
