"""Shared helpers of the -m gpu parity tests (CUDA engine through the C ABI vs the oracle / goldens)."""
import numpy as np


def pack_games(g):
    """Golden board file -> (moves[games, maxplies], colors, counts, first-ply index per game)."""
    game = g["game"]
    ng = int(game.max()) + 1
    counts = np.bincount(game, minlength=ng).astype(np.int32)
    mp = int(counts.max())
    moves = np.zeros((ng, mp), np.int16)
    colors = np.ones((ng, mp), np.uint8)
    start = np.zeros(ng, np.int64)
    i = 0
    for k in range(ng):
        start[k] = i
        moves[k, :counts[k]] = g["pos"][i:i + counts[k]]
        colors[k, :counts[k]] = g["mover"][i:i + counts[k]]
        i += counts[k]
    return moves, colors, counts, start
