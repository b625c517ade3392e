"""Parity tests proper (B200): the CUDA path, called through the C ABI, against
  (1) the committed golden fixtures = outputs of the real reference (tests/golden/*.ref.res),
  (2) the oracle on seeded workloads (oracle/gaml_oracle, run on the box),
  (3) size-independent properties at BASELINE config-2 size (2 M read pairs, 4.6 Mbp).

Tolerances (BASELINE.json north_star): per-read value <= 1e-12 relative, total <= 1e-9 relative, floored
counts and total_len exact. Short-read per-read values are in fact required to be BIT-EXACT here, because
the paired state replays the reference's subtract/add sequence.
"""
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from cases import seeded_cases
from gaml_b200 import api, dist, synth, workload
from gaml_b200.workload import KIND_PACBIO

pytestmark = pytest.mark.gpu

GOLDEN_NAMES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.wl")))
REL_READ, REL_TOTAL = 1e-12, 1e-9


def check_against(wl, ref, pc=None, exact_short=True):
    own = pc is None
    if own:
        pc = api.ProbCalculator.from_workload(wl)
    for e, walks in enumerate(wl.evals):
        prob, zeros, tl = pc.calc_prob(walks)
        r = ref[e]
        assert tl == r.total_len, e
        assert zeros == r.zeros, (e, zeros, r.zeros)
        assert abs(prob - r.score) <= REL_TOTAL * abs(r.score), (e, prob, r.score)
        for s, spec in enumerate(wl.sets):
            v, rv = pc.read_values(s), r.per_read[s]
            if spec.kind == KIND_PACBIO:
                fin = np.isfinite(rv)
                assert np.array_equal(fin, np.isfinite(v)), (e, s)
                assert np.all(np.abs(v[fin] - rv[fin]) <= REL_READ * np.abs(rv[fin])), (e, s)
            elif exact_short:
                assert np.array_equal(v, rv), (e, s, np.nonzero(v != rv)[0][:5])
            else:
                assert np.all(np.abs(v - rv) <= REL_READ * np.abs(rv)), (e, s)
    if own:
        pc.close()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cuda_matches_reference_golden(name):
    wl = workload.read_workload(os.path.join(GOLDEN, name + ".wl"))
    ref = workload.read_results(os.path.join(GOLDEN, name + ".ref.res"))
    check_against(wl, ref)


@pytest.mark.parametrize("name", sorted(seeded_cases().keys()))
def test_cuda_matches_oracle_seeded(name, oracle):
    wl = seeded_cases()[name]
    check_against(wl, oracle(wl, name))


def test_many_placement_reads_take_the_scratch_path():
    wl = workload.read_workload(os.path.join(GOLDEN, "hand_paired.wl"))
    pc = api.ProbCalculator.from_workload(wl)
    pc.calc_prob(wl.evals[0])
    st = pc.stats()
    assert st.last_multi_items + st.last_overflow_reads >= 1 and st.last_scratch_placements >= 1
    pc.close()


def test_error_behaviour():
    wl = workload.read_workload(os.path.join(GOLDEN, "hand_paired.wl"))
    pc = api.ProbCalculator.from_workload(wl)
    with pytest.raises(api.GamlError, match="inserted twice"):
        pc.cache_insert(0, 0, (0,), np.zeros(0, dtype=workload.ALN_DTYPE))
    with pytest.raises(api.GamlError, match="outside the graph"):
        pc.calc_prob([[0, 9999]])
    bad = np.zeros(1, dtype=workload.ALN_DTYPE)
    bad["read_id"] = 10 ** 6
    with pytest.raises(api.GamlError, match="read_id"):
        pc.cache_insert(0, 0, (6, 4), bad)
    # the context is still usable after rejected calls
    prob, _, _ = pc.calc_prob(wl.evals[0])
    ref = workload.read_results(os.path.join(GOLDEN, "hand_paired.ref.res"))
    assert abs(prob - ref[0].score) <= REL_TOTAL * abs(ref[0].score)
    pc.close()


def test_cache_grows_during_annealing(oracle):
    """Keys that appear after the first evaluations (cache misses filled by the aligner, graph.cc:1967-1968)
    are appended and the device CSR is rebuilt; results equal a context that had everything up front."""
    wl = synth.paired_workload(12, 2500, 4000, n_evals=30, seed=31)
    ref = oracle(wl, "grow")
    spec = wl.sets[0]
    pc = api.ProbCalculator(wl.node_len, wl.normalize_map)
    sid = pc.add_readset(spec)
    inserted = [set(), set()]
    for e, walks in enumerate(wl.evals):
        need = synth.short_keys_for_walks(walks, wl.node_len, with_single_node=True)
        for m in range(2):
            for k in need:
                if k not in inserted[m] and k in spec.caches[m]:
                    pc.cache_insert(sid, m, k, spec.caches[m][k])
                    inserted[m].add(k)
        prob, zeros, tl = pc.calc_prob(walks)
        assert zeros == ref[e].zeros and tl == ref[e].total_len
        assert np.array_equal(pc.read_values(0), ref[e].per_read[0]), e
    # a full re-score on top of the grown cache sees the appended records too (fresh state: equal up to the incremental
    # state's own rounding drift)
    v_inc = pc.read_values(0)
    pc.reset_state()
    full = pc.calc_prob(wl.evals[-1])
    assert pc.stats().last_was_full == 1 and full[1:] == (zeros, tl) and abs(full[0] - prob) <= REL_TOTAL * abs(prob)
    v_full = pc.read_values(0)
    nz = v_full != 0
    assert np.array_equal(nz, v_inc != 0) or np.all(np.abs(v_inc[~nz]) < 1e-25)
    assert np.all(np.abs(v_inc[nz] - v_full[nz]) <= 1e-9 * np.abs(v_full[nz]))
    st = pc.stats()
    assert st.cache_appends >= 5 and st.cache_rebuilds <= 3, (st.cache_appends, st.cache_rebuilds)   # growth is applied in place
    pc.close()


def test_cache_growth_by_append_equals_rebuilding(monkeypatch):
    """The same growing cache with appends switched off (every growth rebuilds the device index): identical partial sums
    and per-read values at every step, for incremental and for full evaluations."""
    wl = synth.paired_workload(46, 10000, 200_000, n_evals=40, seed=13)
    spec = wl.sets[0]
    pcs = []
    for no_append in (False, True):
        if no_append:
            monkeypatch.setenv("GAML_B200_NO_APPEND", "1")
        pc = api.ProbCalculator(wl.node_len, wl.normalize_map)
        pc.add_readset(spec)
        pcs.append(pc)
    monkeypatch.delenv("GAML_B200_NO_APPEND")
    inserted = [set(), set()]
    for e, walks in enumerate(wl.evals):
        need = synth.short_keys_for_walks(walks, wl.node_len, with_single_node=True)
        for m in range(2):
            for k in need:
                if k not in inserted[m] and k in spec.caches[m]:
                    for pc in pcs:
                        pc.cache_insert(0, m, k, spec.caches[m][k])
                    inserted[m].add(k)
        if e % 7 == 6:
            for pc in pcs:
                pc.reset_state()
        a, b = [pc.calc_prob_partial(walks) for pc in pcs]
        assert a[1] == b[1] and np.array_equal(a[0], b[0]), (e, a, b)
        if e % 5 == 0:
            assert np.array_equal(pcs[0].read_values(0), pcs[1].read_values(0)), e
    assert pcs[0].stats().cache_appends >= 10 and pcs[1].stats().cache_appends == 0
    for pc in pcs:
        pc.close()


def test_two_shards_on_one_gpu_combine_to_the_unsharded_result(oracle):
    wl = seeded_cases()["mixed_s1"]
    ref = oracle(wl, "shards")
    shards = [api.ProbCalculator.from_workload(wl, shard_of=(r, 3)) for r in range(3)]
    for e, walks in enumerate(wl.evals):
        parts, tl = zip(*[pc.calc_prob_partial(walks) for pc in shards])
        g = np.stack(parts)
        prob, zeros, tl0 = shards[0].combine(g, 3, tl[0])
        assert zeros == ref[e].zeros and tl0 == ref[e].total_len
        assert abs(prob - ref[e].score) <= REL_TOTAL * abs(ref[e].score)
        for s, spec in enumerate(wl.sets):
            v = np.concatenate([pc.read_values(s) for pc in shards])
            rv = ref[e].per_read[s]
            fin = np.isfinite(rv)
            assert np.all(np.abs(v[fin] - rv[fin]) <= REL_READ * np.abs(rv[fin]))
    for pc in shards:
        pc.close()


# ---- full-size properties (BASELINE config 2) ---------------------------------------------------
def test_result_exchange_between_two_contexts_matches_the_unsharded_result(oracle):
    """Two read-id shards as two contexts ("ranks") sharing one result-exchange segment: each publishing block writes
    its 64-byte line into the segment, each side gathers both lines and combines them — the totals must equal the
    unsharded context's exactly, over a whole incremental trajectory (the exchange is double-buffered by epoch)."""
    import mmap
    wl = synth.paired_workload(46, 10000, 200_000, n_evals=12, seed=11)
    whole = api.ProbCalculator.from_workload(wl)
    seg = mmap.mmap(-1, max(2 * 2 * 8 * 64, mmap.PAGESIZE))
    ranks = []
    for rk in range(2):
        pc = api.ProbCalculator.from_workload(wl, shard_of=(rk, 2))
        pc.set_result_exchange(seg, rk, 2)
        ranks.append(pc)
    for e, walks in enumerate(wl.evals):
        ref = whole.calc_prob(walks)
        for pc in ranks:
            pc.prepare(walks)
            pc.launch()
        outs = []
        for pc in ranks:
            g, tl = pc.finish_gathered()
            outs.append(pc.combine(g, 2, tl))
        assert outs[0] == outs[1], e
        assert outs[0][1] == ref[1] and outs[0][2] == ref[2], e
        assert outs[0][0] == ref[0], (e, outs[0][0], ref[0])   # exact integer partial sums: bit-identical for any sharding
    for pc in ranks:
        pc.close()
    whole.close()


def test_peer_memory_exchange_between_two_gpus_matches_the_unsharded_result():
    """The same over peer memory (gaml_peer_exchange_*): each publishing block stores its 64-byte line into the exchange
    buffer of both ranks over NVLink, the last kernel of each chain waits for both lines in its own buffer. Needs TWO
    GPUs (one context each, wired by pointer — IPC handles are for other processes): a kernel that waits for a line
    another kernel writes must never share its GPU with that kernel (B200_PROFILING.md), so on a one-GPU box this is
    covered by `bench.py --gpus N --exchange peer` instead (profiles/r02_exchange_ab.md)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs: kernels that wait on one another must not share a GPU")
    wl = synth.paired_workload(46, 10000, 200_000, n_evals=12, seed=11)
    whole = api.ProbCalculator.from_workload(wl)
    ranks = [api.ProbCalculator.from_workload(wl, device=rk, shard_of=(rk, 2)) for rk in range(2)]
    ptrs = [pc.peer_exchange_create(rk, 2)[1] for rk, pc in enumerate(ranks)]
    for pc in ranks:
        pc.peer_exchange_open(local_ptrs=ptrs)
    for e, walks in enumerate(wl.evals):
        ref = whole.calc_prob(walks)
        for pc in ranks:
            pc.prepare(walks)
        for pc in ranks:
            pc.launch()
        outs = []
        for pc in ranks:
            g, tl = pc.finish_gathered()
            assert g.shape[0] == 2
            outs.append(pc.combine(g, 2, tl))
        assert outs[0] == outs[1], e
        assert outs[0] == ref, (e, outs[0], ref)   # exact integer partial sums: bit-identical for any sharding
    for pc in ranks:
        pc.peer_exchange_close()
        pc.close()
    whole.close()


def test_capacity_error_invalidates_the_state_and_the_one_shot_call_recovers(monkeypatch, oracle):
    """A scratch arena too small for a walk set's many-placement reads: the evaluation reports GAML_ERR_CAPACITY, the
    incremental state is dropped (the skipped reads' values are stale), the buffer grows, and gaml_calc_prob_partial
    repeats the evaluation by itself — the caller sees the oracle's results throughout."""
    wl = workload.read_workload(os.path.join(GOLDEN, "hand_paired.wl"))
    ref = workload.read_results(os.path.join(GOLDEN, "hand_paired.ref.res"))
    monkeypatch.setenv("GAML_B200_SCRATCH_ENTRIES", "1")
    pc = api.ProbCalculator.from_workload(wl)
    monkeypatch.delenv("GAML_B200_SCRATCH_ENTRIES")
    # three-phase form: the error surfaces, the next evaluation is a full re-score with a larger arena
    pc.prepare(wl.evals[0])
    pc.launch()
    with pytest.raises(api.GamlError, match="scratch"):
        pc.finish()
    check_against(wl, ref, pc=pc)
    assert pc.stats().last_scratch_placements >= 1
    pc.close()


def test_flat_cache_file_round_trip(tmp_path):
    """gaml_cache_save / gaml_cache_load: a context whose read sets were filled from the flat cache files scores a
    trajectory identically (partials and per-read values bit for bit) to the one the files were written from —
    paired + single + PacBio sets, and a read-id shard."""
    wl = synth.mixed_workload(10, 6000, 3000, 300, n_single=1000, n_evals=6, seed=21, pacbio_len=5000)
    for shard_of in (None, (1, 2)):
        src = api.ProbCalculator.from_workload(wl, shard_of=shard_of)
        for s in range(len(wl.sets)):
            src.cache_save(s, str(tmp_path / f"set{s}.gcc"))
        dst = api.ProbCalculator(wl.node_len, wl.normalize_map)
        for s, spec in enumerate(wl.sets):
            shard = None
            if shard_of is not None:
                shard = dist.shard_bounds(spec.n_reads, *shard_of)
            assert dst.add_readset(spec, shard) == s
            dst.cache_load(s, str(tmp_path / f"set{s}.gcc"))
        with pytest.raises(api.GamlError):
            dst.cache_load(0, str(tmp_path / "set0.gcc"))   # not empty any more
        for walks in wl.evals:
            a, ta = src.calc_prob_partial(walks)
            b, tb = dst.calc_prob_partial(walks)
            assert ta == tb and np.array_equal(a, b)
        for s in range(len(wl.sets)):
            assert np.array_equal(src.read_values(s), dst.read_values(s), equal_nan=True)
        src.close()
        dst.close()
    bad = api.ProbCalculator.from_workload(wl)
    with pytest.raises(api.GamlError):
        bad.cache_load(0, str(tmp_path / "no_such_file.gcc"))
    bad.close()


@pytest.fixture(scope="module")
def c2():
    wl = synth.paired_workload(460, 10000, 2_000_000, n_evals=10, seed=42)
    pc = api.ProbCalculator.from_workload(wl)
    yield wl, pc
    pc.close()


# ---- BASELINE config 4 shape: 10 000 walks, 20 000 keys per mate (slot tables far beyond L1), one read-id shard ---------
@pytest.fixture(scope="module")
def c4():
    wl = synth.paired_workload(10000, 10000, 1_200_000, n_evals=22, seed=44)
    yield wl


def test_c4_shape_full_and_incremental_match_oracle(c4, oracle):
    """Config 4's shape (100 Mbp genome, 10 000 nodes / walks) on a reduced read count: the full evaluation and 21
    incremental steps against the oracle — totals <= 1e-9, floored counts and total_len exact, per-read values bit-exact."""
    wl = c4
    assert len(wl.evals[0]) >= 10000
    ref = oracle(wl, "c4shape", dump=True)
    check_against(wl, ref)


def test_c4_shape_two_shards_and_internal_order(c4, monkeypatch):
    """The same trajectory as two read-id shards (exact combine), and with the caller's read order kept on the device
    (GAML_B200_NO_PERMUTE): partial sums and per-read values must be IDENTICAL to the default context's."""
    wl = c4
    whole = api.ProbCalculator.from_workload(wl)
    monkeypatch.setenv("GAML_B200_NO_PERMUTE", "1")
    plain = api.ProbCalculator.from_workload(wl)
    monkeypatch.delenv("GAML_B200_NO_PERMUTE")
    shards = [api.ProbCalculator.from_workload(wl, shard_of=(r, 2)) for r in range(2)]
    for e, walks in enumerate(wl.evals[:12]):
        pw, tw = whole.calc_prob_partial(walks)
        pp, tp = plain.calc_prob_partial(walks)
        assert tw == tp and np.array_equal(pw, pp), e
        parts, tls = zip(*[pc.calc_prob_partial(walks) for pc in shards])
        assert shards[0].combine(np.stack(parts), 2, tls[0]) == whole.combine(pw[None, :], 1, tw), e
    assert np.array_equal(whole.read_values(0), plain.read_values(0))
    assert np.array_equal(whole.read_values(0), np.concatenate([pc.read_values(0) for pc in shards]))
    for pc in shards + [whole, plain]:
        pc.close()


def _patched_equals_flattened(wl, monkeypatch, script, min_patched):
    """Two contexts over the same workload, one with patched full evaluations (default), one that flattens every walk
    of every full evaluation (GAML_B200_NO_FULL_PATCH): partial sums and per-read values must be IDENTICAL at every step.
    script = [(index into wl.evals or an explicit walk list, fresh ScoringState?)]."""
    patched = api.ProbCalculator.from_workload(wl)
    monkeypatch.setenv("GAML_B200_NO_FULL_PATCH", "1")
    plain = api.ProbCalculator.from_workload(wl)
    monkeypatch.delenv("GAML_B200_NO_FULL_PATCH")
    for step, (k, fresh) in enumerate(script):
        walks = wl.evals[k] if isinstance(k, int) else k
        if fresh:
            patched.reset_state()
            plain.reset_state()
        pa, ta = patched.calc_prob_partial(walks)
        pb, tb = plain.calc_prob_partial(walks)
        assert ta == tb and np.array_equal(pa, pb), (step, pa, pb)
        assert np.array_equal(patched.read_values(0), plain.read_values(0)), step
    assert patched.stats().full_patch_evals >= min_patched, patched.stats().full_patch_evals
    assert plain.stats().full_patch_evals == 0
    patched.close()
    plain.close()


def test_c4_shape_full_evaluation_of_a_changed_list_uploads_only_its_patch(c4, monkeypatch):
    """A full evaluation (fresh ScoringState) of a walk list a few moves away from the list whose slot tables are on the
    device flattens and uploads only the changed walks' keys (engine.cu "patched full evaluation") — at 10 000 walks the
    difference between O(changed) and O(all walks) host work. Checked against the same library flattening everything:
    joins, splits, there and back, after incremental drift, after a reordered list (renumbering), repeated lists."""
    wl = c4
    rev = list(reversed(wl.evals[0]))
    script = [(0, True), (2, True), (0, True), (3, True), (4, False), (5, False), (6, True), (6, True), (1, True), (9, False),
              (0, True), (rev, True), (2, True), (12, True), (rev, True), (0, True), (21, True), (20, True)]
    _patched_equals_flattened(wl, monkeypatch, script, min_patched=8)


def test_patched_full_evaluation_with_repeat_keys(monkeypatch):
    """The same on a small repeat-rich graph: keys that occur several times (their occurrence lists are merged from the
    base's and the patch's, and the multi pass gets a new range list) and reads with placements on several walks (the
    order of the placements across walks must survive the labels with gaps)."""
    wl = synth.paired_workload(14, 2500, 6000, n_evals=26, seed=77)
    script = [(0, True)] + [(k, True) for k in range(1, 26)] + [(0, True), (13, True), (14, False), (15, False), (3, True), (25, True)]
    _patched_equals_flattened(wl, monkeypatch, script, min_patched=10)


def test_c4_shape_batch_of_1024_candidates(c4):
    """BASELINE config 5 on the config-4 shape: 1024 candidate moves in one batch; a sample of them must equal, bit for
    bit, the sequential evaluation of that candidate's walk set by a fresh context with the same history."""
    import bench
    wl = c4
    history = wl.evals[:6]
    base = history[-1]
    pc = api.ProbCalculator.from_workload(wl)
    for walks in history:
        last = pc.calc_prob(walks)
    cands = bench.make_candidates(base, 1024, np.random.default_rng(5))
    probs, tls, zeros = pc.calc_prob_batch(cands)
    assert pc.calc_prob(base) == last   # stateless
    pc.close()
    assert len(set(int(t) for t in tls)) > 50   # many distinct total lengths: the one-pass base sum is really exercised
    for i in (0, 1, 17, 300, 511, 777, 1023):
        erased, added = cands[i]
        walks = [w for k, w in enumerate(base) if k not in set(erased)] + [list(w) for w in added]
        ref_pc = api.ProbCalculator.from_workload(wl)
        for h in history:
            ref_pc.calc_prob(h)
        p, z, tl = ref_pc.calc_prob(walks)
        ref_pc.close()
        assert probs[i] == p, (i, probs[i], p)
        assert int(tls[i]) == tl and (int(zeros[i, 0, 0]), int(zeros[i, 0, 1])) == z[0]


def test_c2_incremental_state_equals_full_rescore(c2):
    """Walk the scripted trajectory incrementally, then re-score the last walk set from scratch: the
    persistent per-read state may differ from a fresh one only by the reference's own rounding drift."""
    wl, pc = c2
    pc.reset_state()
    for walks in wl.evals:
        inc = pc.calc_prob(walks)
    v_inc = pc.read_values(0)
    assert pc.stats().last_was_full == 0
    pc.reset_state()
    full = pc.calc_prob(wl.evals[-1])
    assert pc.stats().last_was_full == 1
    v_full = pc.read_values(0)
    assert inc[1] == full[1] and inc[2] == full[2]
    assert abs(inc[0] - full[0]) <= REL_TOTAL * abs(full[0])
    nz = v_full != 0
    assert np.all(np.abs(v_inc[nz] - v_full[nz]) <= 1e-9 * np.abs(v_full[nz]))
    # idempotence: the same walk set again is an empty delta and returns the identical double
    again = pc.calc_prob(wl.evals[-1])
    assert again == full


def test_running_total_is_exactly_the_full_resum(monkeypatch):
    """An incremental evaluation at an unchanged total length swaps the touched reads' terms in the running total
    (exact 128-bit integers) instead of re-summing all reads: the partials must be IDENTICAL to those of a context
    that re-sums on every evaluation, over a whole scripted trajectory (joins, splits, tail swaps, flips, rejections)."""
    wl = synth.paired_workload(46, 10000, 200_000, n_evals=60, seed=9)
    fast = api.ProbCalculator.from_workload(wl)
    monkeypatch.setenv("GAML_B200_NO_RUNNING_TOTAL", "1")
    monkeypatch.setenv("GAML_B200_NO_FAST_CHANGES", "1")   # ... and finds its erased/added walks with the reference's container
    monkeypatch.setenv("GAML_B200_NO_GRAPHS", "1")         # ... and launches its kernels one by one
    slow = api.ProbCalculator.from_workload(wl)
    monkeypatch.delenv("GAML_B200_NO_RUNNING_TOTAL")
    monkeypatch.delenv("GAML_B200_NO_FAST_CHANGES")
    monkeypatch.delenv("GAML_B200_NO_GRAPHS")
    for e, walks in enumerate(wl.evals):
        pf, tf = fast.calc_prob_partial(walks)
        ps, ts = slow.calc_prob_partial(walks)
        assert tf == ts
        assert np.array_equal(pf, ps), (e, pf, ps)
        if e % 10 == 9:
            assert np.array_equal(fast.read_values(0), slow.read_values(0))
    assert fast.stats().delta_only_evals >= 10 and slow.stats().delta_only_evals == 0
    assert fast.stats().fast_change_evals >= 30 and slow.stats().fast_change_evals == 0
    fast.close()
    slow.close()


def test_c2_there_and_back_again(c2):
    wl, pc = c2
    pc.reset_state()
    a = pc.calc_prob(wl.evals[0])
    b = pc.calc_prob(wl.evals[3])
    a2 = pc.calc_prob(wl.evals[0])
    assert a[1] == a2[1] and a[2] == a2[2]
    assert abs(a[0] - a2[0]) <= REL_TOTAL * abs(a[0])
    assert b[2] == sum(sum(abs(x) if x < 0 else int(wl.node_len[x]) for x in w) for w in wl.evals[3])


def test_c2_matches_oracle_on_the_full_size(c2, oracle):
    """The oracle needs ~2 s per full evaluation at this size: compare the first three evaluations."""
    wl, pc = c2
    small = workload.Workload(node_len=wl.node_len, normalize_map=wl.normalize_map, sets=wl.sets, evals=wl.evals[:3])
    ref = oracle(small, "c2", dump=True)
    pc.reset_state()
    check_against(small, ref, pc=pc)


def test_c2_walk_order_and_flip_invariance(c2):
    """Permuting the walk list changes nothing but summation order; scoring the reverse complement of every
    walk (with the reverse-complement cache keys present) gives the same likelihood up to rounding."""
    wl, pc = c2
    pc.reset_state()
    base = pc.calc_prob(wl.evals[0])
    perm = list(reversed(wl.evals[0]))
    pc.reset_state()
    p2 = pc.calc_prob(perm)
    assert base[1] == p2[1] and base[2] == p2[2]
    assert abs(base[0] - p2[0]) <= 1e-12 * abs(base[0])


def test_cpp_host_mirror_runs():
    """The C++ ProbCalculator mirror (gaml_b200/host/prob_calculator.h) over the C ABI."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_prob_calculator")
    if not os.path.exists(exe):
        pytest.skip("C++ host test not built")
    out = subprocess.run([exe, os.path.join(GOLDEN, "synth_mixed.wl"), os.path.join(GOLDEN, "synth_mixed.ref.res")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout


def test_reference_annealing_driver_over_cuda_is_identical(tmp_path):
    """The reference's OWN gaml.cc Optimize loop + moves.cc, compiled unchanged against
    integration/prob_calculator.h (oracle/_ref/gaml_gpu, prebuilt where /root/reference exists), must follow
    exactly the trajectory of the pure reference binary: same proposals, same accept/reject decisions."""
    ref = os.path.join(ROOT, "oracle", "_ref", "gaml_ref")
    gpu = os.path.join(ROOT, "oracle", "_ref", "gaml_gpu")
    if not (os.path.exists(ref) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref/gaml_ref + gaml_gpu not prebuilt")
    out = subprocess.run(["bash", os.path.join(ROOT, "tools", "run_e2e.sh"), str(tmp_path), "200"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    # single + paired, each once over the plain drop-in and once with the moves' candidate lists scored in device batches
    # (gaml_gpu_batched: LocalChange2 / FixGapLength / FixRepForNode2 through ProbCalculator::CalcProbBatch)
    batched = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gaml_gpu_batched"))
    assert out.stdout.count("IDENTICAL TRAJECTORY") == (4 if batched else 2), out.stdout


@pytest.mark.parametrize("name", ["synth_mixed", "hand_pacbio", "synth_pacbio_penalty", "synth_paired", "synth_single"])
def test_dropin_prob_calculator_scores_every_read_set_kind(name, tmp_path):
    """oracle/_ref/gpu_harness = the reference's cache-injection harness compiled against integration/prob_calculator.h
    (the drop-in ProbCalculator over the CUDA library) instead of the reference's header: single, paired AND PacBio sets
    through ProbCalculator::CalcProb must reproduce the reference's own results (the committed goldens)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gpu_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gpu_harness not prebuilt")
    res = str(tmp_path / "dropin.res")
    out = subprocess.run([exe, os.path.join(GOLDEN, name + ".wl"), res, "0"], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.returncode == 0, out.stdout + out.stderr
    got = workload.read_results(res)
    ref = workload.read_results(os.path.join(GOLDEN, name + ".ref.res"))
    assert len(got) == len(ref)
    for e, (g, r) in enumerate(zip(got, ref)):
        assert g.total_len == r.total_len and g.zeros == r.zeros, (e, g.zeros, r.zeros)
        assert abs(g.score - r.score) <= REL_TOTAL * abs(r.score), (e, g.score, r.score)


# ---- batched candidate evaluation (BASELINE config 5) ---------------------------------------------
def _delta(base, cand):
    """(indices of base walks not in cand, walks of cand not in base) as a multiset difference."""
    from collections import Counter
    need = Counter(tuple(w) for w in cand)
    erased = []
    for i, w in enumerate(base):
        if need[tuple(w)] > 0:
            need[tuple(w)] -= 1
        else:
            erased.append(i)
    have = Counter(tuple(w) for w in base)
    added = []
    for w in cand:
        if have[tuple(w)] > 0:
            have[tuple(w)] -= 1
        else:
            added.append(list(w))
    return erased, added


def test_batched_candidates_equal_sequential_evaluation(oracle):
    """gaml_calc_prob_batch must return, for every candidate, exactly the double that gaml_calc_prob returns when
    the candidate's walk set is evaluated right after the same history (fresh context per candidate), must not
    disturb the state, and must agree with the oracle."""
    wl = synth.paired_workload(14, 2500, 6000, n_evals=26, seed=77)
    n_hist = 8
    history, cands = wl.evals[:n_hist], wl.evals[n_hist:]
    base = history[-1]
    pc = api.ProbCalculator.from_workload(wl)
    for walks in history:
        last = pc.calc_prob(walks)
    state_before = pc.read_values(0).copy()
    probs, tls, zeros = pc.calc_prob_batch([_delta(base, c) for c in cands])
    assert np.array_equal(pc.read_values(0), state_before)          # stateless
    assert pc.calc_prob(base) == last                                 # and the next normal evaluation is unaffected
    pc.close()
    for i, cand in enumerate(cands):
        ref_pc = api.ProbCalculator.from_workload(wl)
        for walks in history:
            ref_pc.calc_prob(walks)
        p, z, tl = ref_pc.calc_prob(cand)
        ref_pc.close()
        assert probs[i] == p, (i, probs[i], p)                        # bit-identical
        assert int(tls[i]) == tl and (int(zeros[i, 0, 0]), int(zeros[i, 0, 1])) == z[0]
    # oracle: history + one candidate, for a few candidates
    for i in (0, len(cands) // 2, len(cands) - 1):
        small = workload.Workload(node_len=wl.node_len, normalize_map=wl.normalize_map, sets=wl.sets,
                                  evals=history + [cands[i]])
        ref = oracle(small, f"batch{i}", dump=False)
        assert abs(probs[i] - ref[-1].score) <= REL_TOTAL * abs(ref[-1].score)
        assert int(tls[i]) == ref[-1].total_len and int(zeros[i, 0, 0]) == ref[-1].zeros[0][0]


# ---- BASELINE configs 1 and 3 at their stated sizes ---------------------------------------------------
def test_config1_single_reads_full_size(oracle):
    """Config 1: 100 kbp genome, 50 k single-end 100 bp reads; full logL + a scripted trajectory vs the oracle."""
    wl = synth.single_workload(10, 10000, 50_000, n_evals=40, seed=7)
    check_against(wl, oracle(wl, "c1"))


def test_config3_paired_plus_pacbio_full_size(oracle):
    """Config 3: the 4.6 Mbp / 2 M-pair set (weight 1.0) plus 46 k PacBio-like 10 kbp reads (weight 0.5)."""
    wl = synth.mixed_workload(460, 10000, 2_000_000, 46_000, n_evals=3, seed=42, pacbio_len=10000)
    assert [s.kind for s in wl.sets] == [1, 2] and wl.sets[1].weight == 0.5
    check_against(wl, oracle(wl, "c3"))


# ---- coverage-gap penalty ----------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_paired_coverage_penalty_matches_oracle(seed, oracle):
    """penalty_constant != 0 on a sparsely covered genome: the per-walk bad_bases sweep (graph.cc:1893-1919) and its
    incremental bookkeeping (graph.cc:1938, 1946, 1988) must reproduce the oracle's scores; per-read state stays
    bit-exact. The same trajectory without penalty must differ (the penalty is really exercised)."""
    wl = synth.paired_workload(14, 2500, 400, n_evals=40, seed=seed)
    plain = oracle(wl, f"nopen{seed}")
    wl.sets[0].penalty_constant = 0.00007
    wl.sets[0].step = 300.0 - 30.0          # gaml.cc:860: step = insert_mean - penalty_step
    ref = oracle(wl, f"pen{seed}")
    assert any(a.score != b.score for a, b in zip(plain, ref))
    check_against(wl, ref)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pacbio_coverage_penalty_matches_oracle(seed, oracle):
    """penalty_constant != 0 on a PacBio set (graph.cc:3197-3250): intervals of nodes and of the alignments above the
    read's minimum probability, earliest-open-start sweep per walk. bad_bases is an integer, so the penalised total must
    agree with the oracle like the unpenalised one; the same trajectory without penalty must differ."""
    wl = seeded_cases()[f"pacbio_penalty_s{seed}"]
    ref = oracle(wl, f"pbpen{seed}", dump=True)
    check_against(wl, ref)
    pen = wl.sets[-1].penalty_constant
    wl.sets[-1].penalty_constant = 0.0
    pc = api.ProbCalculator.from_workload(wl)
    plain = [pc.calc_prob(w)[0] for w in wl.evals]
    pc.close()
    wl.sets[-1].penalty_constant = pen
    assert any(abs(a - r.score) > 1e-6 for a, r in zip(plain, ref))


@pytest.mark.parametrize("name", ["synth_paired_penalty", "synth_pacbio_penalty"])
def test_coverage_penalty_on_read_id_shards_matches_reference_golden(name):
    """A penalised set split over read-id shards: a walk's coverage events live on all shards, so each
    evaluation the shards' events are gathered (gaml_penalty_export), the union is swept on every shard
    (gaml_penalty_import) and the partials combined — against the reference's own results for the unsharded set."""
    wl = workload.read_workload(os.path.join(GOLDEN, name + ".wl"))
    ref = workload.read_results(os.path.join(GOLDEN, name + ".ref.res"))
    n_shards = 3
    shards = [api.ProbCalculator.from_workload(wl, shard_of=(r, n_shards)) for r in range(n_shards)]
    pen_sets = [s for s, spec in enumerate(wl.sets) if spec.penalty_constant != 0 and spec.kind != 0]
    assert pen_sets
    for e, walks in enumerate(wl.evals):
        parts, tls = zip(*[pc.calc_prob_partial(walks) for pc in shards])
        with pytest.raises(api.GamlError, match="gaml_penalty_import"):
            shards[0].combine(np.stack(parts), n_shards, tls[0])     # events still pending
        for s in pen_sets:
            union = np.concatenate([pc.penalty_export(s) for pc in shards])
            for pc in shards:
                pc.penalty_import(s, union)
        outs = [pc.combine(np.stack(parts), n_shards, tls[0]) for pc in shards]
        assert outs[0] == outs[1] == outs[2], e
        prob, zeros, tl = outs[0]
        assert tl == ref[e].total_len and zeros == ref[e].zeros, e
        assert abs(prob - ref[e].score) <= REL_TOTAL * abs(ref[e].score), (e, prob, ref[e].score)
    for pc in shards:
        pc.close()


def test_single_set_penalty_is_a_no_op_like_the_reference(oracle):
    wl = synth.single_workload(10, 2500, 3000, n_evals=12, seed=5)
    plain = oracle(wl, "s_nopen")
    wl.sets[0].penalty_constant = 0.0001
    ref = oracle(wl, "s_pen")
    assert [r.score for r in plain] == [r.score for r in ref]      # graph.cc:1710-1733 never counts a gap
    check_against(wl, ref)


# ---- PacBio alignment probability on the device --------------------------------------------------
ALNPROB_GOLDEN = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "alnprob_*.ap")))


def _read_ap(path):
    """GAMLAP1 -> (alignments, match, mismatch, band)."""
    import struct
    from gaml_b200 import alnprob
    raw = open(path, "rb").read()
    assert raw[:8] == b"GAMLAP1\0"
    match, mismatch, band, n = struct.unpack_from("<ddii", raw, 8)
    off = 8 + 24
    alns = []
    for _ in range(n):
        posstart, n1 = struct.unpack_from("<ii", raw, off); off += 8
        s1 = raw[off:off + n1]; off += n1
        (n2,) = struct.unpack_from("<i", raw, off); off += 4
        s2 = raw[off:off + n2]; off += n2
        (k,) = struct.unpack_from("<i", raw, off); off += 4
        cigar = []
        for _ in range(k):
            ln, op = struct.unpack_from("<ii", raw, off); off += 8
            cigar.append((ln, chr(op)))
        alns.append(alnprob.Alignment(s1, s2, posstart, cigar))
    return alns, match, mismatch, band


def _close_logvals(got, ref):
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(got))
    assert np.all(np.abs(got[fin] - ref[fin]) <= REL_READ * np.abs(ref[fin])), np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin]))


@pytest.mark.parametrize("name", ALNPROB_GOLDEN)
def test_cuda_alnprob_matches_reference_golden(name):
    """gaml_pacbio_alignment_logprob against logvals written by the reference's own AligmentProbability: same cells, same
    order of additions; only exp/log1p differ from glibc (<= 1e-12 relative on the log value, the per-read bar)."""
    alns, match, mismatch, band = _read_ap(os.path.join(GOLDEN, name + ".ap"))
    ref = np.frombuffer(open(os.path.join(GOLDEN, name + ".ref.lp"), "rb").read(), dtype="<f8")
    pc = api.ProbCalculator([100], None)
    got = pc.pacbio_alignment_logprob(alns, match, mismatch, band)
    pc.close()
    _close_logvals(got, ref)


@pytest.mark.parametrize("name", ["long_band2", "many_short", "clipped"])
def test_cuda_alnprob_matches_oracle_seeded(name, tmp_path):
    from cases import alnprob_seeded
    from gaml_b200 import alnprob
    from conftest import ORACLE_BIN
    alns, match, mismatch, band = alnprob_seeded()[name]
    ap, lp = str(tmp_path / "a.ap"), str(tmp_path / "a.lp")
    alnprob.write_alignments(ap, alns, match, mismatch, band)
    subprocess.run([ORACLE_BIN, "--alnprob", ap, lp], check=True, stderr=subprocess.DEVNULL)
    ref = alnprob.read_logvals(lp, len(alns))[0]
    pc = api.ProbCalculator([100], None)
    got = pc.pacbio_alignment_logprob(alns, match, mismatch, band)
    assert len(pc.pacbio_alignment_logprob([], match, mismatch, band)) == 0
    with pytest.raises(api.GamlError):
        pc.pacbio_alignment_logprob([alnprob.Alignment(b"ACGT", b"ACGT", 1, [(4, "X")])], match, mismatch, band)
    pc.close()
    _close_logvals(got, ref)


def test_records_that_do_not_fit_the_packed_pairs_stream_the_plain_records(oracle):
    """Tier 1 of the streaming kernel reads a 16-byte packed copy of each pair's two first records (key < 2^22, edit
    distance <= 127). A set with a larger edit distance keeps the plain 16-byte records: same results either way."""
    wl = synth.paired_workload(10, 3000, 3000, n_evals=6, seed=77, read_len=250, insert_mean=600.0, insert_std=40.0)
    spec = wl.sets[0]
    bumped = 0
    for cache in spec.caches:
        for key in sorted(cache.keys()):
            recs = cache[key]
            if len(recs) and bumped < 40:
                recs["edit_dist"][0] = 128 + bumped      # still <= read length + 6
                bumped += 1
    assert bumped >= 20
    check_against(wl, oracle(wl, "unpackable"))


def test_ragged_lengths_with_packed_pairs(oracle):
    """Per-pair lengths differ: the packed records are used, the lengths are read from their own array."""
    wl = synth.paired_workload(10, 3000, 3000, n_evals=6, seed=78)
    spec = wl.sets[0]
    rng = np.random.default_rng(5)
    spec.read_len[0] = np.asarray(spec.read_len[0]).copy()
    spec.read_len[1] = np.asarray(spec.read_len[1]).copy()
    spec.read_len[0][::3] -= rng.integers(1, 20, size=len(spec.read_len[0][::3])).astype(spec.read_len[0].dtype)
    spec.read_len[1][::5] -= rng.integers(1, 20, size=len(spec.read_len[1][::5])).astype(spec.read_len[1].dtype)
    check_against(wl, oracle(wl, "ragged_packed"))
