"""CPU tests that pin the oracle (oracle/gaml_oracle.cc) against the REAL reference:

  * tests/golden/*.ref.res were written by oracle/_ref/ref_harness, i.e. by the reference's own compiled
    graph.cc / prob_calculator.h (tests/golden/make_golden.py). They travel to the GPU box.
  * when oracle/_ref/ref_harness is present (this container, or prebuilt on the box) larger seeded
    workloads are run through both binaries.
The bar is bit-exact: scores, total_len, floored counts and every per-read value.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, ORACLE_BIN, REF_HARNESS, run_scorer
from cases import seeded_cases
from gaml_b200 import workload

GOLDEN_NAMES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.wl")))


def assert_bit_exact(got, ref):
    assert len(got) == len(ref)
    for e, (g, r) in enumerate(zip(got, ref)):
        assert g.total_len == r.total_len, e
        assert g.zeros == r.zeros, e
        assert g.score == r.score or (np.isnan(g.score) and np.isnan(r.score)), (e, g.score, r.score)
        assert len(g.per_read) == len(r.per_read)
        for s, (a, b) in enumerate(zip(g.per_read, r.per_read)):
            assert np.array_equal(a, b), (e, s)


def test_golden_present():
    assert len(GOLDEN_NAMES) >= 6


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_reference_golden(name, tmp_path):
    ref = workload.read_results(os.path.join(GOLDEN, name + ".ref.res"))
    got = run_scorer(ORACLE_BIN, os.path.join(GOLDEN, name + ".wl"), tmp_path, name)
    assert_bit_exact(got, ref)


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("name", sorted(seeded_cases().keys()))
def test_oracle_matches_reference_live(name, tmp_path):
    wl = seeded_cases()[name]
    ref = run_scorer(REF_HARNESS, wl, tmp_path, name)
    got = run_scorer(ORACLE_BIN, wl, tmp_path, name)
    assert_bit_exact(got, ref)


def test_golden_cover_edge_cases():
    """The fixtures really contain the edge cases the docstrings promise."""
    hp = workload.read_workload(os.path.join(GOLDEN, "hand_paired.wl"))
    assert [] in hp.evals                                   # empty assembly
    assert any(w.count([0]) == 2 for w in hp.evals)         # duplicated walk (multiset diff)
    assert any(any(x < 0 for x in walk) for ws in hp.evals for walk in ws)   # gap
    recs = hp.sets[0].caches[0][(2,)]
    assert (recs["read_id"] == 6).sum() > 8                 # many-placement read (scratch path)
    res = workload.read_results(os.path.join(GOLDEN, "hand_paired.ref.res"))
    assert any(z[0][0] > 0 for z in (r.zeros for r in res))  # floored reads exist
    pb = workload.read_workload(os.path.join(GOLDEN, "hand_pacbio.wl"))
    assert (pb.normalize_map != np.arange(len(pb.normalize_map))).any()
    rp = workload.read_results(os.path.join(GOLDEN, "hand_pacbio.ref.res"))
    assert np.isneginf(rp[0].per_read[0]).any()             # reads with no alignment: -inf LSE identity


# ---- PacBio alignment probability (PacbioReadSet::AligmentProbability, graph.cc:2175-2297) -------------------------
def _run_alnprob(binary, ap_path, n, tmp_path, tag):
    import subprocess
    from gaml_b200 import alnprob
    out = os.path.join(str(tmp_path), tag + ".lp")
    subprocess.run([binary, "--alnprob", ap_path, out], check=True, stderr=subprocess.DEVNULL)
    return alnprob.read_logvals(out, n)[0]


ALNPROB_GOLDEN = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "alnprob_*.ap")))


def test_alnprob_golden_present():
    assert len(ALNPROB_GOLDEN) >= 3


@pytest.mark.parametrize("name", ALNPROB_GOLDEN)
def test_oracle_alnprob_matches_reference_golden(name, tmp_path):
    """The committed logvals came out of the reference's own AligmentProbability (ref_harness --alnprob)."""
    ref = np.frombuffer(open(os.path.join(GOLDEN, name + ".ref.lp"), "rb").read(), dtype="<f8")
    got = _run_alnprob(ORACLE_BIN, os.path.join(GOLDEN, name + ".ap"), len(ref), tmp_path, name)
    assert np.array_equal(got, ref), np.nonzero(got != ref)[0][:5]
    assert np.isinf(ref).sum() < len(ref)


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("name", ["long_band2", "many_short", "clipped"])
def test_oracle_alnprob_matches_reference_live(name, tmp_path):
    from cases import alnprob_seeded
    from gaml_b200 import alnprob
    alns, match, mismatch, band = alnprob_seeded()[name]
    ap = os.path.join(str(tmp_path), name + ".ap")
    alnprob.write_alignments(ap, alns, match, mismatch, band)
    ref = _run_alnprob(REF_HARNESS, ap, len(alns), tmp_path, name + "_ref")
    got = _run_alnprob(ORACLE_BIN, ap, len(alns), tmp_path, name + "_ora")
    assert np.array_equal(got, ref)
