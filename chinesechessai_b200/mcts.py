"""Batched MCTS on the GPU: one tree per game, all games advance wave by wave.

Device-side counterpart of ``MCTS.search`` (self_play.py:89-154): per wave ONE
``xq_mcts_select`` launch walks every tree to its network leaf, ONE evaluator call scores all
leaves (the only dense contraction — a single batched ``ChessNet.forward``), ONE
``xq_mcts_backup`` launch expands and backs up.  No host synchronisation inside a search.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from . import _lib
from ._lib import BOARD_STRIDE, MAX_MOVES, check
from .engine import (_ptr, _stream, bias_residual_relu, encode_planes, encode_planes_nhwc16,
                     policy_priors, stem_lookup)

WAVE = 8  # self_play.py:101


class HashEvaluator:
    """Deterministic stand-in for the network (the function is specified in DESIGN.md);
    used by parity tests and by search-only benchmarks."""

    def __init__(self, flat: bool = False):
        self.flat = flat
        self.lib = _lib.load()

    def __call__(self, leaf_board, leaf_player, leaf_moves, leaf_n):
        n = leaf_board.shape[0]
        pri = torch.empty((n, MAX_MOVES), dtype=torch.float32, device=leaf_board.device)
        val = torch.empty((n,), dtype=torch.float64, device=leaf_board.device)
        with torch.cuda.device(leaf_board.device):
            check(self.lib.xq_hash_eval(_ptr(leaf_board), leaf_board.stride(0), _ptr(leaf_player),
                                        _ptr(leaf_moves), _ptr(leaf_n), 1 if self.flat else 0,
                                        _ptr(pri), _ptr(val), n, _stream()))
        return pri, val


class _FoldedNet(torch.nn.Module):
    """Inference copy of ChessNet for the search: eval-mode BatchNorm folded into the preceding
    conv (exact algebra, torch.nn.utils.fusion), weights cast to ``dtype``, channels-last,
    conv+bias+ReLU(+residual) issued as ONE cuDNN call each (torch.cudnn_convolution_relu /
    _add_relu) instead of conv + separate bias-add / ReLU / add kernels, and the policy head
    padded from 8100 to 8192 outputs so the GEMM takes an aligned tensor-core path (the prior
    kernel reads the padded rows through its row stride).  Library calls only — no custom GEMM."""

    POLICY_PAD = 8192

    def __init__(self, net: torch.nn.Module, dtype: torch.dtype):
        super().__init__()
        import copy
        from torch.nn.utils.fusion import fuse_conv_bn_eval
        n = copy.deepcopy(net).eval()
        self.stem = fuse_conv_bn_eval(n.conv1, n.bn1)
        self.in_channels = self.stem.in_channels
        if dtype == torch.bfloat16 and self.stem.in_channels == 15:
            # 16 input channels (one zero plane): 32-byte pixels in channels-last, so cuDNN needs
            # no padding pass and the planes come straight from xq_encode_planes_nhwc16
            st = self.stem
            padded = torch.nn.Conv2d(16, st.out_channels, st.kernel_size, st.stride, st.padding,
                                     bias=True, device=st.weight.device)
            with torch.no_grad():
                padded.weight.zero_()
                padded.weight[:, :15].copy_(st.weight)
                padded.bias.copy_(st.bias)
            self.stem = padded
            self.in_channels = 16
        self.blocks = torch.nn.ModuleList(
            torch.nn.ModuleList([fuse_conv_bn_eval(b.conv1, b.bn1), fuse_conv_bn_eval(b.conv2, b.bn2)])
            for b in n.res_blocks)
        self.policy_conv = fuse_conv_bn_eval(n.policy_conv, n.policy_bn)
        self.value_conv = fuse_conv_bn_eval(n.value_conv, n.value_bn)
        fc = n.policy_fc
        self.policy_fc = torch.nn.Linear(fc.in_features, self.POLICY_PAD, device=fc.weight.device)
        with torch.no_grad():
            self.policy_fc.weight.zero_()
            self.policy_fc.bias.zero_()
            self.policy_fc.weight[:fc.out_features].copy_(fc.weight)
            self.policy_fc.bias[:fc.out_features].copy_(fc.bias)
        self.value_fc1, self.value_fc2 = n.value_fc1, n.value_fc2
        # The reference flattens NCHW (.view): feature index = c*90 + square.  The activations here
        # are channels-last, so the two FC layers get their input columns re-ordered once to
        # square*C + c and read the conv outputs as they lie in memory (no transpose pass).
        with torch.no_grad():
            for fc_layer, c_in in ((self.policy_fc, self.policy_conv.out_channels),
                                   (self.value_fc1, self.value_conv.out_channels)):
                w = fc_layer.weight
                fc_layer.weight.copy_(w.view(w.shape[0], c_in, 90).permute(0, 2, 1).reshape(w.shape[0], -1))
        self.to(dtype=dtype, memory_format=torch.channels_last)
        self.fused = False
        self.own_epilogue = dtype == torch.bfloat16
        # encode_board + conv1/bn1/ReLU as a lookup (xq_stem_lookup_bf16): table[tap][plane][c_out]
        # from the same bf16 weights the cuDNN stem uses
        self.stem_table = None
        st = self.stem
        if (dtype == torch.bfloat16 and self.in_channels == 16 and tuple(st.kernel_size) == (3, 3) and
                tuple(st.stride) == (1, 1) and tuple(st.padding) == (1, 1) and st.out_channels == 128 and
                st.weight.is_cuda):
            with torch.no_grad():
                self.stem_table = st.weight.detach().permute(2, 3, 1, 0).reshape(9, 16, st.out_channels).contiguous()
                # bias per (side to move, square): plane 14 is all ones when red moves, so its
                # taps that fall on the board add a square-dependent constant
                b = st.bias.detach().float()
                w14 = self.stem_table[:, 14].float()                       # [9, C]
                on = torch.tensor([[1.0 if 0 <= sq // 9 + tap // 3 - 1 <= 9 and 0 <= sq % 9 + tap % 3 - 1 <= 8
                                    else 0.0 for tap in range(9)] for sq in range(90)], device=b.device)
                self.stem_bias = torch.stack([b[None, :].expand(90, -1), b[None, :] + on @ w14]).contiguous()
        if next(self.parameters()).is_cuda:
            try:  # probe the fused cuDNN entry points once
                x = torch.zeros((2, self.in_channels, 10, 9), dtype=dtype, device=fc.weight.device).contiguous(
                    memory_format=torch.channels_last)
                self.fused = True
                self.forward(x)
                torch.cuda.synchronize()
            except Exception:
                self.fused = False

    @staticmethod
    def _cr(conv, x):
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding,
                                            conv.dilation, conv.groups)

    def forward_boards(self, board, player):
        """The network on raw positions (int8 boards + side to move): the stem runs as the fused
        lookup kernel, the rest as in forward()."""
        if self.stem_table is None or not self.fused:
            return self.forward(encode_planes_nhwc16(board, player) if self.in_channels == 16 else
                                encode_planes(board, player, dtype=self.stem.weight.dtype).contiguous(
                                    memory_format=torch.channels_last))
        return self._trunk(stem_lookup(board, player, self.stem_table, self.stem_bias))

    def forward(self, x):
        if self.fused:
            return self._trunk(self._cr(self.stem, x))
        return self._trunk(torch.relu(self.stem(x)))

    def _trunk(self, x):
        if self.fused:
            for c1, c2 in self.blocks:
                y = self._cr(c1, x)
                if self.own_epilogue:  # plain conv + one fused HBM pass (our kernel)
                    z = torch.nn.functional.conv2d(y, c2.weight, None, c2.stride, c2.padding)
                    x = bias_residual_relu(z, x, c2.bias)
                else:
                    x = torch.cudnn_convolution_add_relu(y, c2.weight, x, 1.0, c2.bias, c2.stride,
                                                         c2.padding, c2.dilation, c2.groups)
            p = self._cr(self.policy_conv, x)
            v = self._cr(self.value_conv, x)
        else:
            for c1, c2 in self.blocks:
                x = torch.relu(c2(torch.relu(c1(x))) + x)
            p = torch.relu(self.policy_conv(x))
            v = torch.relu(self.value_conv(x))
        nb = p.shape[0]
        p = self.policy_fc(p.permute(0, 2, 3, 1).reshape(nb, -1))  # a view of channels-last memory
        v = torch.tanh(self.value_fc2(torch.relu(self.value_fc1(v.permute(0, 2, 3, 1).reshape(nb, -1)))))
        return p, v


class NetEvaluator:
    """encode_board -> ChessNet.forward -> gather+softmax, all on device
    (neural_network.py:96-126 without the per-sample D2H of :120-124).

    ``dtype=float32`` runs the module as given (the reference's precision).  ``tf32`` = True /
    False forces TF32 tensor cores on / off for cuDNN and cuBLAS during the forward; None leaves
    PyTorch's switches alone, which is what the reference itself gets on a GPU (convolutions in
    TF32 by default, matmuls in strict fp32).  A lower-precision dtype — or float32 with
    ``tf32=True``, whose rounding is coarser than the BN folding's — builds a folded
    inference copy (BN folded, ``dtype`` weights, channels-last, conv+bias+ReLU(+residual) as
    one cuDNN call each; ``folded`` overrides the choice).  The copy is rebuilt whenever
    the module's weights have changed since it was made — optimizer steps, ``load_state_dict``
    and the NCCL weight broadcast all bump the tensors' version counters, which is what
    ``version`` watches — so an evaluator kept across training iterations never plays with stale
    weights.  The module must be in eval mode (the reference's workers force it,
    self_play.py:339,346): BatchNorm batch statistics must not leak into the search."""

    def __init__(self, net: torch.nn.Module, dtype: torch.dtype = torch.float32,
                 tf32: Optional[bool] = None, folded: Optional[bool] = None):
        self.net = net
        self.dtype = dtype
        self.tf32 = tf32      # None: PyTorch's own switches (cuDNN TF32 on, cuBLAS off by default)
        self.folded = (dtype != torch.float32 or tf32 is True) if folded is None else bool(folded)
        self._fast = None
        self._seen = None     # weight fingerprint the folded copy / captured graphs belong to
        self._version = 0

    def _fingerprint(self):
        return tuple(t._version for t in self.net.state_dict(keep_vars=True).values())

    @property
    def version(self) -> int:
        """Changes whenever the weights do; captured CUDA graphs of a search are keyed on it."""
        fp = self._fingerprint()
        if fp != self._seen:
            self._seen = fp
            self._fast = None
            self._version += 1
        return self._version

    def refresh(self) -> None:
        """Force a rebuild of the folded copy (kept for callers that modify weights through
        ``.data`` or other paths that bypass the version counters)."""
        self._seen = None

    @torch.no_grad()
    def __call__(self, leaf_board, leaf_player, leaf_moves, leaf_n):
        if self.net.training:
            raise RuntimeError("NetEvaluator: the network is in train() mode; call network.eval() before "
                               "self-play (BatchNorm would use and update batch statistics)")
        _ = self.version
        if self.folded:
            with _tf32(self.tf32):
                if self._fast is None:
                    self._fast = _FoldedNet(self.net, self.dtype)
                logits, value = self._fast.forward_boards(leaf_board, leaf_player)
        else:
            planes = encode_planes(leaf_board, leaf_player, dtype=self.dtype)
            with _tf32(self.tf32):
                logits, value = self.net(planes)
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        pri = policy_priors(logits, leaf_moves, leaf_n)
        return pri, value.reshape(-1).float().contiguous()


class _tf32:
    """Scope for the TF32 switches of cuDNN convolutions and cuBLAS matmuls."""

    def __init__(self, on: Optional[bool]):
        self.on = on          # None: leave PyTorch's switches alone

    def __enter__(self):
        self.prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        if self.on is not None:
            torch.backends.cudnn.allow_tf32 = bool(self.on)
            torch.backends.cuda.matmul.allow_tf32 = bool(self.on)

    def __exit__(self, *exc):
        if self.on is not None:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


class BatchedMCTS:
    def __init__(self, n_games: int, num_simulations: int, device: Optional[torch.device] = None):
        self.lib = _lib.load()
        _lib.require_device()
        self.n = int(n_games)
        self.num_simulations = int(num_simulations)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = self.device
        self.tree_bytes = int(self.lib.xq_mcts_tree_bytes(self.num_simulations))
        self.trees = torch.empty((max(self.n, 1) * self.tree_bytes,), dtype=torch.uint8, device=d)
        self.leaf_board = torch.zeros((self.n, BOARD_STRIDE), dtype=torch.int8, device=d)
        self.leaf_player = torch.ones((self.n,), dtype=torch.int8, device=d)
        self.leaf_moves = torch.zeros((self.n, MAX_MOVES), dtype=torch.int16, device=d)
        self.leaf_n = torch.zeros((self.n,), dtype=torch.int16, device=d)
        self.leaf_mult = torch.zeros((self.n,), dtype=torch.int16, device=d)
        self.root_moves = torch.zeros((self.n, MAX_MOVES), dtype=torch.int16, device=d)
        self.root_visits = torch.zeros((self.n, MAX_MOVES), dtype=torch.int32, device=d)
        self.root_n = torch.zeros((self.n,), dtype=torch.int16, device=d)
        self._cbuf = None
        self.rows_evaluated = 0   # rows the evaluator has been run on (n per wave without compaction)

    # -- the three kernels ---------------------------------------------------------------
    def init(self, board: torch.Tensor, meta: torch.Tensor, active: Optional[torch.Tensor] = None):
        with torch.cuda.device(self.device):
            check(self.lib.xq_mcts_init(_ptr(self.trees), self.num_simulations, _ptr(board), _ptr(meta),
                                        _ptr(active), self.n, _stream()))

    def select(self, wave_size: int):
        with torch.cuda.device(self.device):
            check(self.lib.xq_mcts_select(_ptr(self.trees), self.num_simulations, wave_size,
                                          _ptr(self.leaf_board), _ptr(self.leaf_player),
                                          _ptr(self.leaf_moves), _ptr(self.leaf_n),
                                          _ptr(self.leaf_mult), self.n, _stream()))

    def backup(self, priors: torch.Tensor, values: torch.Tensor, values_per_game: int = 1):
        assert priors.dtype == torch.float32 and priors.is_contiguous()
        assert values.dtype in (torch.float32, torch.float64) and values.is_contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.xq_mcts_backup(_ptr(self.trees), self.num_simulations, _ptr(self.leaf_moves),
                                          _ptr(self.leaf_n), _ptr(priors), _ptr(values),
                                          1 if values.dtype == torch.float32 else 0,
                                          values_per_game, self.n, _stream()))

    def visits(self):
        with torch.cuda.device(self.device):
            check(self.lib.xq_mcts_root_visits(_ptr(self.trees), self.num_simulations,
                                               _ptr(self.root_moves), _ptr(self.root_visits),
                                               _ptr(self.root_n), self.n, _stream()))
        return self.root_moves, self.root_visits, self.root_n

    # -- leaf compaction (ragged batches) ------------------------------------------------------
    def _compact_buffers(self):
        if self._cbuf is None:
            d, n = self.device, self.n
            self._cbuf = dict(
                idx=torch.zeros((n,), dtype=torch.int32, device=d),
                row=torch.zeros((n,), dtype=torch.int32, device=d),
                count=torch.zeros((1,), dtype=torch.int32, device=d),
                board=torch.zeros((n, BOARD_STRIDE), dtype=torch.int8, device=d),
                player=torch.ones((n,), dtype=torch.int8, device=d),
                moves=torch.zeros((n, MAX_MOVES), dtype=torch.int16, device=d),
                n=torch.zeros((n,), dtype=torch.int16, device=d))
        return self._cbuf

    def bucket(self, bound: int) -> int:
        """Rows to evaluate for at most ``bound`` live leaves: rounded up to one of <= 16 sizes,
        so that cuDNN / cuBLAS / the allocator see a handful of shapes, not one per ply."""
        q = max(64, self.n // 16)
        return self.n if bound >= self.n else max(q, min(self.n, -(-int(bound) // q) * q))

    def evaluate_and_backup(self, evaluator: Callable, rows: Optional[int] = None) -> None:
        """The evaluator call and xq_mcts_backup of one wave.  ``rows`` (an upper bound of the
        number of games with a network leaf, from ``bucket()``) < n compacts the leaves first:
        the reference never sends a finished game or a terminal leaf to the network
        (self_play.py:126-139), a batch would otherwise run the forward on all n rows."""
        if rows is None or rows >= self.n:
            self.rows_evaluated += self.n
            priors, values = evaluator(self.leaf_board, self.leaf_player, self.leaf_moves, self.leaf_n)
            self.backup(priors, values)
            return
        c = self._compact_buffers()
        with torch.cuda.device(self.device):
            st = _stream()
            check(self.lib.xq_compact_leaves(_ptr(self.leaf_n), self.n, _ptr(c["idx"]), _ptr(c["row"]),
                                             _ptr(c["count"]), st))
            check(self.lib.xq_gather_leaves(_ptr(c["idx"]), _ptr(c["count"]), rows, _ptr(self.leaf_board),
                                            _ptr(self.leaf_player), _ptr(self.leaf_moves), _ptr(self.leaf_n),
                                            _ptr(c["board"]), _ptr(c["player"]), _ptr(c["moves"]),
                                            _ptr(c["n"]), st))
        self.rows_evaluated += rows
        priors, values = evaluator(c["board"][:rows], c["player"][:rows], c["moves"][:rows], c["n"][:rows])
        assert priors.dtype == torch.float32 and priors.is_contiguous() and values.is_contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.xq_mcts_backup_rows(_ptr(self.trees), self.num_simulations, _ptr(self.leaf_moves),
                                               _ptr(self.leaf_n), _ptr(priors), _ptr(values),
                                               1 if values.dtype == torch.float32 else 0, _ptr(c["row"]),
                                               self.n, _stream()))

    # -- MCTS.search for the whole batch ---------------------------------------------------
    def search(self, board: torch.Tensor, meta: torch.Tensor, evaluator: Callable,
               active: Optional[torch.Tensor] = None, rows: Optional[int] = None
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Returns (moves int16[n,128], visit counts int32[n,128], n_children int16[n]); rows are
        the root's children in ``get_legal_moves()`` order, like the dict of self_play.py:151-154.
        ``rows``: see ``evaluate_and_backup`` (the caller guarantees that at most ``rows`` games
        reach a network leaf in any wave, e.g. the number of games still running)."""
        self.init(board, meta, active)
        for start in range(0, self.num_simulations, WAVE):
            self.select(min(WAVE, self.num_simulations - start))
            self.evaluate_and_backup(evaluator, rows)
        return self.visits()
