"""Self-play records -> training data against the reference's own generator output.

tests/golden/rldata_9.npz was written by the reference's nn/data_generator.py on the golden self-play SGFs
(tests/golden/make_golden.py gen_rldata).  The host half (sampling, symmetry, policy / value targets) is checked here
on the CPU; the device half (positions replayed on the board pool + feature-plane kernel) in the -m gpu test."""
import os

import numpy as np
import pytest


def _samples(golden_dir):
    from tamago_b200.nn import data_generator as dg
    g = np.load(os.path.join(golden_dir, "rldata_9.npz"))
    sp = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    texts = [str(sp["sgf"][i]) for i in g["order"]]
    np.random.seed(int(g["seed"]))
    return g, dg.sample_positions(texts, 9, literal=True)


def test_sampling_policy_and_value_targets_match_reference(golden_dir):
    g, samples = _samples(golden_dir)
    assert len(samples) == len(g["value"]) == 24
    assert np.array_equal(np.array([s[4] for s in samples], np.int32), g["value"])
    pol = np.array([s[3] for s in samples])
    assert pol.shape == g["policy"].shape and np.array_equal(pol, g["policy"])


def test_symmetry_maps_are_permutations():
    from tamago_b200.nn.data_generator import symmetry_maps
    for n in (9, 19):
        m = symmetry_maps(n)
        assert (np.sort(m, axis=1) == np.arange(n * n)).all() and (m[0] == np.arange(n * n)).all()
        assert len({tuple(r) for r in m}) == 8


def test_sgf_reader_roundtrip(golden_dir):
    from tamago_b200.sgf.reader import SGFReader
    sp = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    for k, text in enumerate(sp["sgf"]):
        r = SGFReader(str(text), 9, literal=True)
        want = sp["moves"][sp["moves_off"][k]:sp["moves_off"][k + 1]]
        assert r.get_moves() == [int(x) for x in want] and r.size == 9
        assert r.get_comment(0).split(" ")[0] == "82"


@pytest.mark.gpu
def test_generated_npz_matches_reference(golden_dir, tmp_path):
    from tamago_b200.nn import data_generator as dg
    g = np.load(os.path.join(golden_dir, "rldata_9.npz"))
    sp = np.load(os.path.join(golden_dir, "selfplay_9.npz"))
    kdir = tmp_path / "kifu"
    kdir.mkdir()
    paths = []
    for i in g["order"]:
        p = kdir / f"{int(i)}.sgf"
        p.write_text(str(sp["sgf"][int(i)]), encoding="utf-8")
        paths.append(str(p))
    old = dg.BATCH_SIZE
    dg.BATCH_SIZE = 8                       # the golden was written with the same patch (24 samples < 256)
    try:
        np.random.seed(int(g["seed"]))
        written = dg.generate_reinforcement_learning_data(str(tmp_path), [str(kdir)], 9, kifu_list=paths)
    finally:
        dg.BATCH_SIZE = old
    assert len(written) == 1
    d = np.load(written[0])
    assert np.array_equal(d["input"], g["input"]) and d["input"].dtype == np.float32
    assert np.array_equal(d["policy"], g["policy"]) and np.array_equal(d["value"], g["value"])
    assert int(d["kifu_count"]) == int(g["kifu_count"])
