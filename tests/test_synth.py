"""CPU: the synthetic workload generators are deterministic and have the stated shape."""
import numpy as np


def test_random_pairs_shape_and_determinism():
    from sequencealigning_b200 import synth
    a = synth.random_pairs(2000, 150, 0.05, True, seed=1)
    b = synth.random_pairs(2000, 150, 0.05, True, seed=1)
    assert np.array_equal(a.residues, b.residues) and np.array_equal(a.d_len, b.d_len)
    assert set(np.unique(a.residues).tolist()) <= set(b"ACGT")
    assert (a.q_len == 150).all() and 140 < a.d_len.mean() < 160 and a.d_len.std() > 0.5
    s = synth.random_pairs(2000, 150, 0.05, False, seed=1)
    assert (s.d_len == 150).all()
    q = np.stack([np.frombuffer(s.query(p), np.uint8) for p in range(200)])
    d = np.stack([np.frombuffer(s.db(p), np.uint8) for p in range(200)])
    assert 0.03 < (q != d).mean() < 0.07  # 5 % substitutions, each one a real change
    # an insertion-free, deletion-free pair of the indel set is identical in length
    assert a.cells == int((a.q_len.astype(np.int64) * a.d_len).sum())


def test_config1_is_one_query_many_db():
    from sequencealigning_b200 import synth
    b = synth.config1(50)
    assert b.n_pairs == 50 and (b.q_off == 0).all() and (b.q_len == 150).all()
