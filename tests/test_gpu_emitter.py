"""SURVEY 8f-1, the direct route: training samples from the device record ring (tg_emit_samples) must be bit-equal to what
the reference's generate_reinforcement_learning_data (nn/data_generator.py:89-149) wrote for the same games and the same
numpy stream (tests/golden/rldata_9.npz) -- without SGF text, parsing or a host replay -- and the samples can be gathered
across ranks over NCCL without leaving the device."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _play_golden_games(tb, g, sample_cap=64):
    size, seed, visits = int(g["size"]), int(g["seed"]), int(g["visits"])
    ng = len(g["sgf"])
    e = tb.Engine(board_size=size, games=ng, max_visits=visits, superko=True, evaluator=tb.EVAL_HASHNET, seed=seed,
                  record_ring=True, sample_cap=sample_cap)
    e.set_zobrist(g["zobrist"])
    e.reset(game_ids=np.arange(ng), never_resign=g["never_resign"])
    n_moves = np.zeros(ng, np.int32)
    done = np.zeros(ng, bool)
    for _ in range(2 * size * size + 2):
        r = e.genmove(mode=tb.MODE_SH, visits=visits, play=True)
        newly = (r["finished"] != 0) & ~done
        n_moves[newly] = r["n_moves"][newly]
        done |= newly
        if done.all():
            break
    assert done.all()
    return e, n_moves


def test_ring_emitter_is_bit_equal_to_the_reference_generator(golden_dir):
    import tamago_b200 as tb
    from tamago_b200.nn.data_generator import draw_samples, save_samples_npz
    g = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    want = np.load(os.path.join(golden_dir, "rldata_9.npz"))
    e, n_moves = _play_golden_games(tb, g)
    for k in range(len(n_moves)):
        assert n_moves[k] == g["moves_off"][k + 1] - g["moves_off"][k]
    np.random.seed(int(want["seed"]))                       # the generator's numpy stream (make_golden.py gen_rldata)
    order = [int(k) for k in want["order"]]
    plies, syms = np.zeros((len(order), 8), np.int32), np.zeros((len(order), 8), np.int32)
    for j, k in enumerate(order):                           # games in the order the reference visited the shuffled files
        plies[j], syms[j] = draw_samples(int(n_moves[k]))
    assert e.emit_samples(order, plies, syms) == len(want["value"])
    inp, pol, val = e.read_samples(round_like_sgf=True)
    assert np.array_equal(inp, want["input"]) and inp.dtype == want["input"].dtype
    assert np.array_equal(val, want["value"])
    assert np.array_equal(pol, want["policy"]), np.abs(pol - want["policy"]).max()
    raw = e.read_samples(round_like_sgf=False)[1]           # the device values before the "%.3e" round trip of the SGF comment
    assert not np.array_equal(raw, pol) and np.allclose(raw, pol, rtol=6e-4, atol=0)
    # zero-copy device views and the npz writer
    ti, tp, tv = e.sample_tensors()
    assert ti.is_cuda and np.array_equal(ti.cpu().numpy(), inp) and np.array_equal(tv.cpu().numpy(), val)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        assert save_samples_npz(e, os.path.join(d, "rl_data_0"), kifu_count=4) == 24
        z = np.load(os.path.join(d, "rl_data_0.npz"))
        assert np.array_equal(z["policy"], want["policy"]) and int(z["kifu_count"]) == 4 and z["value"].dtype == np.int32
    with pytest.raises(Exception):
        e.emit_samples([0], np.full((1, 8), -1, np.int32) + np.arange(8) * 0 + 500, syms[:1])      # ply outside the game
    e.clear_samples()
    assert e.sample_count == 0
    e.close()


def test_pool_emits_samples_while_playing(tmp_path):
    """SelfPlayPool(sample_cap=...): finished games feed the device sample arrays as they end; every sample is a legal
    training row (one-hot stone planes, targets that sum to ~1, labels in {0,1,2}) and the SGF side output still appears."""
    import tamago_b200 as tb
    from tamago_b200.selfplay.worker import SelfPlayPool
    np.random.seed(3)
    pool = SelfPlayPool(str(tmp_path), 9, 16, 12, iter(range(1, 25)), evaluator=tb.EVAL_HASHNET, seed=5, sample_cap=24 * 8)
    pool.start()
    while pool.active.any():
        pool.step()
    assert pool.files == 24 and pool.samples == 24 * 8
    inp, pol, val = pool.eng.read_samples()
    assert ((inp[:, 0] + inp[:, 1] + inp[:, 2]) == 1).all() and set(np.unique(val)) <= {0, 1, 2}
    np.testing.assert_allclose(pol.sum(axis=1), 1.0, atol=2e-3)
    assert len(os.listdir(tmp_path)) == 24
    pool.close()
