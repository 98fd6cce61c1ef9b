"""Policy/value network of the reference (neural_network.py:12-71, :172-187) — same architecture
and state_dict keys, so the reference's checkpoints (trainer.py:438-443) load unchanged.  The
forward pass stays plain PyTorch (the only dense contraction on the path); board encoding and the
logits -> move-prior gather/softmax run in the CUDA kernels xq_encode_planes / xq_policy_priors.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .engine import encode_planes, pack_move, policy_priors

Move = Tuple[int, int, int, int]


def _pack_checked(legal_moves) -> List[int]:
    """Policy indices (from*90+to, neural_network.py:160) of a caller-supplied move list.  The
    kernels index the logits with them, so off-board moves and lists longer than the engine's
    128-move capacity are refused here instead of reaching the device."""
    if len(legal_moves) > 128:
        raise ValueError(f"{len(legal_moves)} legal moves exceed the engine's capacity of 128")
    out = []
    for m in legal_moves:
        fr, fc, tr, tc = (int(x) for x in m)
        if not (0 <= fr < 10 and 0 <= tr < 10 and 0 <= fc < 9 and 0 <= tc < 9):
            raise ValueError(f"move {tuple(m)} is off the 10x9 board")
        out.append(pack_move((fr, fc, tr, tc)))
    return out


class ResidualBlock(nn.Module):
    def __init__(self, num_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(num_channels, num_channels, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(num_channels)
        self.conv2 = nn.Conv2d(num_channels, num_channels, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(num_channels)

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(self.bn2(self.conv2(y)) + x)


class ChessNet(nn.Module):
    """15x10x9 planes -> (8100 move logits, tanh value); 24,634,141 parameters."""

    def __init__(self, num_channels: int = 128):
        super().__init__()
        self.conv1 = nn.Conv2d(15, num_channels, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(num_channels)
        self.res_blocks = nn.ModuleList(ResidualBlock(num_channels) for _ in range(4))
        self.policy_conv = nn.Conv2d(num_channels, 32, 1)
        self.policy_bn = nn.BatchNorm2d(32)
        self.policy_fc = nn.Linear(32 * 90, 90 * 90)
        self.value_conv = nn.Conv2d(num_channels, 8, 1)
        self.value_bn = nn.BatchNorm2d(8)
        self.value_fc1 = nn.Linear(8 * 90, 128)
        self.value_fc2 = nn.Linear(128, 1)

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(x)))
        for blk in self.res_blocks:
            x = blk(x)
        p = torch.relu(self.policy_bn(self.policy_conv(x))).flatten(1)
        p = self.policy_fc(p)
        v = torch.relu(self.value_bn(self.value_conv(x))).flatten(1)
        v = torch.tanh(self.value_fc2(torch.relu(self.value_fc1(v))))
        return p, v

    # -- reference call surface (neural_network.py:73-169) ---------------------------------
    def _device(self) -> torch.device:
        d = next(self.parameters()).device
        if d.type != "cuda":
            raise RuntimeError("ChessNet's encode/prior kernels need the module on a CUDA device "
                               "(no CPU fallback)")
        return d

    def encode_board(self, board, current_player) -> np.ndarray:
        d = self._device()
        b = torch.from_numpy(np.ascontiguousarray(board, dtype=np.int8).reshape(1, 90)).to(d)
        p = torch.tensor([1 if current_player == 1 else -1], dtype=torch.int8, device=d)
        return encode_planes(b, p)[0].cpu().numpy()

    @torch.no_grad()
    def predict_batch(self, boards_and_players_and_moves: Sequence) -> List[Tuple[Dict[Move, np.float32], float]]:
        items = list(boards_and_players_and_moves)
        if not items:
            return []
        d = self._device()
        n = len(items)
        boards = np.stack([np.asarray(b, dtype=np.int8).reshape(90) for b, _, _ in items])
        players = np.array([1 if p == 1 else -1 for _, p, _ in items], np.int8)
        moves = np.zeros((n, 128), np.int16)
        counts = np.zeros(n, np.int16)
        for i, (_, _, lm) in enumerate(items):
            counts[i] = len(lm)
            moves[i, :len(lm)] = _pack_checked(lm)
        tb, tp = torch.from_numpy(boards).to(d), torch.from_numpy(players).to(d)
        tm, tn = torch.from_numpy(moves).to(d), torch.from_numpy(counts).to(d)
        logits, values = self.forward(encode_planes(tb, tp))
        pri = policy_priors(logits.float().contiguous(), tm, tn).cpu().numpy()
        vals = values.reshape(-1).float().cpu().numpy()
        return [({tuple(m): pri[i, j] for j, m in enumerate(lm)}, float(vals[i]))
                for i, (_, _, lm) in enumerate(items)]

    def predict(self, board, current_player, legal_moves):
        return self.predict_batch([(board, current_player, legal_moves)])[0]

    def _logits_to_move_probs(self, logits, legal_moves) -> Dict[Move, np.float32]:
        if len(legal_moves) == 0:
            return {}
        d = self._device()
        lg = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32).reshape(1, -1)).to(d)
        mv = torch.zeros((1, 128), dtype=torch.int16, device=d)
        mv[0, :len(legal_moves)] = torch.tensor(_pack_checked(legal_moves), dtype=torch.int16)
        cnt = torch.tensor([len(legal_moves)], dtype=torch.int16, device=d)
        pri = policy_priors(lg, mv, cnt)[0].cpu().numpy()
        return {tuple(m): pri[j] for j, m in enumerate(legal_moves)}


def test_network():
    """``python main.py test`` calls this (main.py:177): one forward pass and the parameter count."""
    print("测试神经网络...")
    net = ChessNet().to("cuda")
    x = torch.randn(4, 15, 10, 9, device="cuda")
    policy, value = net(x)
    print(f"输入形状: {x.shape}  策略输出形状: {policy.shape}  价值输出形状: {value.shape}")
    print(f"价值范围: {value.min().item():.2f} ~ {value.max().item():.2f}")
    print(f"总参数量: {sum(p.numel() for p in net.parameters()):,}")
