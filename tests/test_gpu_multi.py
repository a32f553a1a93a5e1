"""GPU: ONE engine over several devices (sa_engine_create_multi).  One sa_align_batch call shards
the pair list of /root/reference/src/main.rs:61-62 over the devices and returns everything in
input order; the results must be bit-identical to the single-device call and to the oracle.

A device id may be listed twice, so the whole multi-device front (planning, worker threads, pool
regions, LPT gather/scatter) is exercised on a one-GPU box too; with >= 2 GPUs the same tests
also run over all of them."""
import numpy as np
import pytest

from tests.util import mutate, random_pair_list, random_seq

pytestmark = pytest.mark.gpu


def _batch(pairs):
    from sequencealigning_b200 import PairBatch
    return PairBatch.from_pairs(pairs)


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(n)))
    return sets


def _same(a, b, what):
    assert np.array_equal(a.score, b.score), what
    assert np.array_equal(a.status, b.status), what
    assert np.array_equal(a.cigar_len, b.cigar_len), what
    # offsets: monotone in p, each pair's words inside the pool; the words themselves are equal
    off = b.cigar_off.astype(np.int64)
    assert (np.diff(off) >= 0).all(), what
    for p in range(a.score.size):
        assert a.cigar_of(p) == b.cigar_of(p), (what, p)


def _oracle_check(oracle, batch, res, what):
    stride = int((batch.q_len.astype(np.int64) + batch.d_len).max()) + 1
    ref = oracle.affine_batch(batch.residues, batch.q_off, batch.q_len, batch.d_off, batch.d_len, cigar_stride=stride, n_threads=8)
    assert np.array_equal(ref.score, res.score), what
    assert np.array_equal(ref.status, res.status), what
    assert np.array_equal(ref.cigar_len, res.cigar_len), what
    for p in range(0, batch.n_pairs, max(1, batch.n_pairs // 500)):
        assert list(ref.cigar_pool[p, :ref.cigar_len[p]]) == res.cigar_of(p), (what, p)


@pytest.mark.parametrize("devices", _device_sets())
def test_contiguous_shards_match_single_device(engine, oracle, devices):
    from sequencealigning_b200 import Engine
    b = _batch(random_pair_list(11, 6000, 0, 260))
    single = engine.align(b)
    with Engine(devices=devices) as multi:
        assert multi.device_count == len(devices)
        r = multi.align(b)
        shards = multi.shards()
        tim = multi.timing()
    _same(single, r, f"devices {devices}")
    _oracle_check(oracle, b, r, f"devices {devices}")
    assert len(shards) == len(devices) and all(s["contiguous"] == 1 for s in shards)
    assert sum(s["pairs"] for s in shards) == b.n_pairs and sum(s["cells"] for s in shards) == b.cells
    assert [s["first_pair"] for s in shards] == list(np.cumsum([0] + [s["pairs"] for s in shards[:-1]]))
    cells = np.array([s["cells"] for s in shards], np.float64)
    assert cells.max() / cells.mean() < 1.02
    assert tim["cells"] == b.cells and tim["kernel_launches"] >= 3 * len(devices)


@pytest.mark.parametrize("devices", _device_sets())
def test_lpt_shards_for_few_uneven_pairs(engine, oracle, devices):
    """No contiguous cut balances these: the call gathers LPT index sets per device and scatters
    the results (and the CIGAR words) back into input order."""
    import random
    from sequencealigning_b200 import Engine
    rng = random.Random(5)
    pairs = []
    for n in [1500, 20, 1400, 35, 700, 50, 650, 10, 3, 0, 2500, 90]:   # 2500: a long pair (general kernel)
        q = random_seq(rng, n, b"ACGT")
        pairs.append((q, mutate(rng, q, 0.08, True, b"ACGT")))
    b = _batch(pairs)
    single = engine.align(b)
    with Engine(devices=devices) as multi:
        r = multi.align(b)
        shards = multi.shards()
    assert all(s["contiguous"] == 0 for s in shards)
    _same(single, r, f"lpt devices {devices}")
    _oracle_check(oracle, b, r, f"lpt devices {devices}")
    # here the pool is rebuilt in input order: offsets are the plain scan
    exp = np.zeros(b.n_pairs, np.uint64)
    exp[1:] = np.cumsum(r.cigar_len[:-1], dtype=np.uint64)
    assert np.array_equal(exp, r.cigar_off) and int(r.cigar_len.sum()) == r.cigar.size


def test_pool_regions_overflow_and_retry(oracle):
    """A pool that is too small for one device's region: SA_E_CIGAR_CAPACITY with a capacity that
    fits in cigar_used; the retry (what Engine.align does) succeeds."""
    from sequencealigning_b200 import Engine, EngineError, _capi
    b = _batch(random_pair_list(13, 3000, 50, 200, rates=(0.3,)))
    with Engine(devices=[0, 0]) as multi:
        with pytest.raises(EngineError) as ei:
            multi.align(b, cigar_capacity=2000)
        assert ei.value.code == _capi.E_CIGAR_CAPACITY
        r = multi.align(b)      # default capacity, grows on demand
        _oracle_check(oracle, b, r, "after retry")
        r0 = multi.align(b, cigar=False)
        assert np.array_equal(r0.score, r.score) and np.array_equal(r0.status, r.status)


def test_other_algorithms_and_modes_through_the_multi_engine(engine):
    from sequencealigning_b200 import (ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD, MODE_LOCAL, NOT_IMPLEMENTED, Engine,
                                       EngineError)
    b = _batch(random_pair_list(17, 1200, 1, 120, alphabet=b"ACGT"))
    with Engine(devices=[0, 0]) as multi:
        for algo in (ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD):
            _same(engine.align(b, algo=algo), multi.align(b, algo=algo), f"algo {algo}")
        r = multi.align(b, mode=MODE_LOCAL)
        assert (r.status == NOT_IMPLEMENTED).all()
        with pytest.raises(EngineError) as ei:
            multi.upload(b)
        assert ei.value.code == -5   # resident batches are single-device
        assert multi.align(_batch([])).score.size == 0


def test_config3_shape_sharded(engine, oracle):
    """BASELINE.json configs[2] in miniature: 250 bp pairs from one host list through one call."""
    import torch
    from sequencealigning_b200 import Engine, synth
    b = synth.random_pairs(40000, 250, 0.05, True, seed=synth.SEEDS["config3"])
    n = max(2, torch.cuda.device_count())
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() >= 2 else [0, 0]
    with Engine(devices=devices) as multi:
        for batch in (b, b.packed()):
            r = multi.align(batch)
            assert np.array_equal(r.score, engine.align(batch, cigar=False).score)
            _oracle_check(oracle, b, r, f"{n} devices, packing {batch.packing}")


@pytest.mark.parametrize("devices", _device_sets())
def test_linear_local_mode_through_the_multi_engine(engine, devices):
    """Local mode (needleman_wunsch.rs:88-89, :107-111) sharded: scores, end cells and CIGARs come back in input
    order, both on the contiguous plan and on the LPT gather/scatter plan."""
    import random
    from sequencealigning_b200 import ALGO_NW_LINEAR, MODE_LOCAL, Engine
    rng = random.Random(23)
    many = _batch(random_pair_list(29, 3000, 0, 200, alphabet=b"ACGT", unrelated=0.4))
    uneven = []
    for n in [900, 20, 800, 35, 400, 50, 10, 3, 0, 1200, 60]:
        q = random_seq(rng, n, b"ACGT")
        uneven.append((random_seq(rng, 40, b"ACGT") + q, mutate(rng, q, 0.1, True, b"ACGT") + random_seq(rng, 25, b"ACGT")))
    with Engine(devices=devices) as multi:
        for b, contiguous in ((many, 1), (_batch(uneven), 0)):
            single = engine.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
            r = multi.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
            assert all(s["contiguous"] == contiguous for s in multi.shards())
            _same(single, r, f"local, devices {devices}")
            assert np.array_equal(single.end1, r.end1) and np.array_equal(single.end2, r.end2)


def test_sampled_plan_of_a_very_large_list(engine):
    """Shards of >= 512 Ki pairs are cut from sampled block sums (multi.cu plan_contiguous): same results as one
    device, exact per-shard cell counts, balance within 1 %."""
    from sequencealigning_b200 import Engine, synth
    b = synth.random_pairs(1_100_000, 24, 0.1, True, seed=99)
    single = engine.align(b)
    with Engine(devices=[0, 0]) as multi:
        r = multi.align(b)
        shards = multi.shards()
    assert np.array_equal(single.score, r.score) and np.array_equal(single.status, r.status)
    assert np.array_equal(single.cigar_len, r.cigar_len)
    for p in range(0, b.n_pairs, 997):   # (the pool of a multi-device call is one region per device)
        assert single.cigar_of(p) == r.cigar_of(p), p
    assert sum(s["pairs"] for s in shards) == b.n_pairs and sum(s["cells"] for s in shards) == b.cells
    cells = np.array([s["cells"] for s in shards], np.float64)
    assert cells.max() / cells.mean() < 1.01
