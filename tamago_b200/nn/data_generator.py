"""Self-play records -> training data (nn/data_generator.py:16-33, 89-149; nn/feature.py:10-57, 80-102).

Same sampling (8 random plies per game, one of the 8 symmetries each, numpy / random global streams consumed in the
reference's order), same npz layout (`input [n,6,N,N] f32`, `policy [n,N*N+1] f64`, `value [n] i32`, `kifu_count`).
The positions are not replayed in Python: all sampled (game, ply) prefixes are played in one batch on the device board
pool and expanded by the feature-plane kernel (tg_play + tg_planes); the symmetry is a host-side index permutation.
"""
import glob
import os
import random

import numpy as np

from ..board.constant import PASS, GTP_X_COORDINATE
from ..sgf.reader import SGFReader

BATCH_SIZE = 256                    # learning_param.py:11
DATA_SET_SIZE = BATCH_SIZE * 4000   # learning_param.py:31


def symmetry_maps(n):
    """sym_map of go_board.py:74-104 restricted to raster indices: perm[sym][i] = raster index of the point that
    supplies output point i under symmetry sym."""
    w = n + 2
    maps = np.zeros((8, n * n), np.int64)
    for y in range(1, n + 1):
        for x in range(1, n + 1):
            i = (y - 1) * n + (x - 1)
            src = [(x, y), (w - (x + 1), y), (x, w - (y + 1)), (w - (x + 1), w - (y + 1)),
                   (y, x), (y, w - (x + 1)), (w - (y + 1), x), (w - (y + 1), w - (x + 1))]
            for s, (sx, sy) in enumerate(src):
                maps[s, i] = (sy - 1) * n + (sx - 1)
    return maps


def draw_samples(n_moves):
    """data_generator.py:123-124: the (plies, symmetries) one game contributes, drawn from numpy's global stream in the
    reference's order.  Returns (plies[8] ascending, -1 padded; syms[8])."""
    target_index = sorted(np.random.permutation(np.arange(n_moves))[:8])
    sym_index_list = np.random.permutation(np.arange(8))
    plies = np.full(8, -1, np.int32)
    plies[:len(target_index)] = target_index
    return plies, sym_index_list.astype(np.int32)


def emit_from_ring(engine, finished_slots, n_moves):
    """SURVEY 8f-1, the direct path: samples of the finished games go from the engine's device record ring to its
    device-resident sample arrays (tg_emit_samples) -- no SGF text, no parser, no host replay, symmetry applied on the
    device.  Call between collect() and the reset that recycles the slots.  Returns the new sample count."""
    plies = np.zeros((len(finished_slots), 8), np.int32)
    syms = np.zeros((len(finished_slots), 8), np.int32)
    for j, nm in enumerate(n_moves):
        plies[j], syms[j] = draw_samples(int(nm))
    return engine.emit_samples(finished_slots, plies, syms)


def save_samples_npz(engine, path, kifu_count, first=0, n=None):
    """rl_data_<k>.npz (data_generator.py:16-33) from the device sample arrays, bit-equal to the SGF route"""
    inp, pol, val = engine.read_samples(first, n, round_like_sgf=True)
    np.savez_compressed(path, input=inp, policy=pol, value=val, kifu_count=np.array(kifu_count))
    return len(val)


def _rl_target(comment, n, perm):
    """generate_rl_target_data (feature.py:80-102): improved policy from the "k pos:prob ..." comment, 1e-18 elsewhere."""
    flat = np.full(n * n, 1e-18, np.float64)
    p_pass = 1e-18
    for datum in comment.split(" ")[1:]:
        if not datum:
            continue
        pos, target = datum.split(":")
        if pos.upper() == "PASS":
            p_pass = float(target)
        else:
            x = GTP_X_COORDINATE.index(pos.upper()[0]) - 1
            y = n - int(pos[1:])
            flat[y * n + x] = float(target)
    return np.append(flat[perm], p_pass)


def _save_data(save_file_path, input_data, policy_data, value_data, kifu_counter):
    """data_generator.py:16-33"""
    np.savez_compressed(save_file_path, input=np.array(input_data[0:DATA_SET_SIZE]), policy=np.array(policy_data[0:DATA_SET_SIZE]),
                        value=np.array(value_data[0:DATA_SET_SIZE], dtype=np.int32), kifu_count=np.array(kifu_counter))


def sample_positions(kifu_list, board_size, literal=False):
    """The host half: which (game, ply, symmetry) are used, with their policy / value targets (data_generator.py:117-137)."""
    maps = symmetry_maps(board_size)
    samples = []                                       # (moves prefix, colours, to_move, sym, policy, value, kifu ordinal)
    for ordinal, kifu in enumerate(kifu_list):
        sgf = SGFReader(kifu, board_size, literal=literal)
        value_label = sgf.get_value_label()
        target_index = sorted(np.random.permutation(np.arange(sgf.get_n_moves()))[:8])
        sym_index_list = np.random.permutation(np.arange(8))
        sym_index = 0
        color = 1
        moves = sgf.get_moves()
        for i, pos in enumerate(moves):
            if i in target_index:
                sym = int(sym_index_list[sym_index])
                samples.append((moves[:i], color, sym, _rl_target(sgf.get_comment(i), board_size, maps[sym]), value_label, ordinal))
                sym_index += 1
            color = 3 - color                          # the reference alternates colours regardless of the SGF colour field
            value_label = 2 - value_label
    return samples


def planes_for_samples(samples, board_size, device=0, chunk=8192):
    """The device half: replay every sampled prefix on the board pool and run the feature-plane kernel."""
    from ..engine import Engine, EVAL_HASHNET
    maps = symmetry_maps(board_size)
    out = np.zeros((len(samples), 6, board_size, board_size), np.float32)
    for c0 in range(0, len(samples), chunk):
        part = samples[c0:c0 + chunk]
        eng = Engine(board_size=board_size, games=len(part), max_visits=2, superko=False, evaluator=EVAL_HASHNET, device=device)
        width = max(1, max(len(s[0]) for s in part))
        mv = np.zeros((len(part), width), np.int16)
        col = np.ones((len(part), width), np.uint8)
        cnt = np.zeros(len(part), np.int32)
        for k, s in enumerate(part):
            mv[k, :len(s[0])] = s[0]
            col[k, :len(s[0])] = [1 + (j % 2) for j in range(len(s[0]))]
            cnt[k] = len(s[0])
        eng.play(mv, cnt, col)
        eng.set_to_move(np.array([s[1] for s in part], np.int32))
        pl = eng.planes().reshape(len(part), 6, board_size * board_size)
        for k, s in enumerate(part):
            out[c0 + k] = pl[k][:, maps[s[2]]].reshape(6, board_size, board_size)
        eng.close()
    return out


def generate_reinforcement_learning_data(program_dir, kifu_dir_list, board_size=9, kifu_list=None, device=0):
    """data_generator.py:89-149 (same file naming: <program_dir>/data/rl_data_<k>.npz)."""
    if kifu_list is None:
        kifu_list = []
        for kifu_dir in kifu_dir_list:
            kifu_list.extend(glob.glob(os.path.join(kifu_dir, "*.sgf")))
        random.shuffle(kifu_list)
    samples = sample_positions(kifu_list, board_size)
    planes = planes_for_samples(samples, board_size, device)
    os.makedirs(os.path.join(program_dir, "data"), exist_ok=True)
    input_data, policy_data, value_data = [], [], []
    kifu_counter, data_counter, k = 1, 0, 0
    written = []
    for ordinal in range(len(kifu_list)):
        while k < len(samples) and samples[k][5] == ordinal:
            input_data.append(planes[k]); policy_data.append(samples[k][3]); value_data.append(samples[k][4]); k += 1
        if len(value_data) >= DATA_SET_SIZE:
            path = os.path.join(program_dir, "data", f"rl_data_{data_counter}")
            _save_data(path, input_data, policy_data, value_data, kifu_counter)
            written.append(path + ".npz")
            input_data, policy_data, value_data = input_data[DATA_SET_SIZE:], policy_data[DATA_SET_SIZE:], value_data[DATA_SET_SIZE:]
            kifu_counter = 1
            data_counter += 1
        kifu_counter += 1
    n_batches = len(value_data) // BATCH_SIZE
    if n_batches > 0:
        path = os.path.join(program_dir, "data", f"rl_data_{data_counter}")
        _save_data(path, input_data[0:n_batches * BATCH_SIZE], policy_data[0:n_batches * BATCH_SIZE],
                   value_data[0:n_batches * BATCH_SIZE], kifu_counter)
        written.append(path + ".npz")
    return written
