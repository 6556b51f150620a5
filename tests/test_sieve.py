"""UAP / CLK1-6 discovery: the oracle's restatement of
btbb_uap_from_header against the unmodified reference (where oracle/_ref exists) and against the
committed fixture generated from it; the CUDA path against the oracle (-m gpu)."""
import json
import os

import numpy as np
import pytest

import util
from util import B

CASES = {                         # name -> sieve_case() arguments
    "clean": dict(n_slots=500, ber=0.0),
    "ber_0.5pct": dict(n_slots=700, ber=0.005, seed=7),
    "ber_2pct": dict(n_slots=900, ber=0.02, seed=11, n_laps=6),
    "clk_step_3": dict(n_slots=400, ber=0.002, seed=5, clk_step=3),
    "hv1_only_no_crc": dict(n_slots=600, n_laps=4, ber=0.003, seed=13, mix=("ID", "HV1")),
    "hv1_noisy": dict(n_slots=800, n_laps=3, ber=0.03, seed=17, mix=("HV1",)),
    "incoherent_resets": dict(n_slots=900, stride=2000, n_laps=3, ber=0.001, seed=19, mix=("HV1", "ID"), coherent=False),
    "wrong_clock_rate": dict(n_slots=600, n_laps=3, ber=0.0, seed=23, mix=("HV1", "DM1"), clk_step=64),
    "stuck_1100_packets": dict(n_slots=60, n_laps=3, ber=0.0, seed=29, mix=("HV1",), repeat_first=1100),
}
FIXTURE = os.path.join(util.GOLDEN, "sieve.json")


def _digest(st, rv):
    return util.digest(st), util.digest(rv)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_sieve_matches_fixture_and_reference(name):
    """The fixture holds digests of the reference's own btbb_uap_from_header results."""
    O = util.oracle()
    stream, pkts, gs, laps, truth = util.sieve_case(**CASES[name])
    st, rv = util.sieve_run(O, "orc", stream, pkts, gs)
    fx = json.load(open(FIXTURE))[name]
    assert [len(pkts), len(gs) - 1] == fx["shape"]
    assert list(_digest(st, rv)) == fx["sha256"], name
    if util.have_ref():
        st_r, rv_r = util.sieve_run(util.ref(), "ref", stream, pkts, gs)
        assert st.tobytes() == st_r.tobytes() and rv.tobytes() == rv_r.tobytes()


def test_oracle_sieve_finds_the_planted_piconets():
    O = util.oracle()
    stream, pkts, gs, laps, truth = util.sieve_case(**CASES["clean"])
    st, rv = util.sieve_run(O, "orc", stream, pkts, gs)
    found = 0
    for g, lap in enumerate(laps):
        if st[g]["flags"] & B.F_UAP_VALID:
            uap, clk0 = truth[lap]
            # CLK1-6 of a packet = CLKN + clk_offset (mod 64); CLKN of slot 0 is 0
            assert st[g]["uap"] == uap and (st[g]["clk_offset"] & 63) == clk0, hex(lap)
            found += 1
    assert found >= len(laps) - 1
    assert (rv == 1).sum() == found and (rv == B.SIEVE_NOT_CALLED).sum() > 0


def test_oracle_sieve_state_carries_across_calls():
    """Feeding a capture in two halves gives the same result as feeding it at once."""
    O = util.oracle()
    stream, pkts, gs, laps, truth = util.sieve_case(**CASES["ber_0.5pct"])
    st_all, rv_all = util.sieve_run(O, "orc", stream, pkts, gs)
    half = [(int(gs[g]) + int(gs[g + 1])) // 2 for g in range(len(gs) - 1)]
    first = np.concatenate([pkts[gs[g]:half[g]] for g in range(len(half))])
    second = np.concatenate([pkts[half[g]:gs[g + 1]] for g in range(len(half))])
    gs1 = np.cumsum([0] + [half[g] - int(gs[g]) for g in range(len(half))]).astype(np.int64)
    gs2 = np.cumsum([0] + [int(gs[g + 1]) - half[g] for g in range(len(half))]).astype(np.int64)
    st1, rv1 = util.sieve_run(O, "orc", stream, first, gs1)
    st2, rv2 = util.sieve_run(O, "orc", stream, second, gs2, states=st1)
    assert st2.tobytes() == st_all.tobytes()
    got = np.concatenate([np.concatenate([rv1[gs1[g]:gs1[g + 1]], rv2[gs2[g]:gs2[g + 1]]]) for g in range(len(half))])
    assert got.tobytes() == rv_all.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_sieve_matches_oracle_and_fixture(gpu_ctx2, orc, name):
    stream, pkts, gs, laps, truth = util.sieve_case(**CASES[name])
    want_st, want_rv = util.sieve_run(orc, "orc", stream, pkts, gs)
    st0 = np.zeros(len(gs) - 1, dtype=B.SIEVE_DTYPE)
    got_st, got_rv = gpu_ctx2.uap_sieve_host(stream, pkts, gs, st0)
    assert got_rv.tobytes() == want_rv.tobytes()
    assert got_st.tobytes() == want_st.tobytes()
    assert list(_digest(got_st, got_rv)) == json.load(open(FIXTURE))[name]["sha256"]


@pytest.mark.gpu
def test_gpu_sieve_in_two_calls_and_empty_groups(gpu_ctx2, orc):
    stream, pkts, gs, laps, truth = util.sieve_case(**CASES["ber_0.5pct"])
    # an empty group in the middle and at the end must leave their states untouched
    gs_e = np.concatenate([gs[:3], gs[2:3], gs[3:], gs[-1:]]).astype(np.int64)
    want_st, want_rv = util.sieve_run(orc, "orc", stream, pkts, gs_e)
    st0 = np.zeros(len(gs_e) - 1, dtype=B.SIEVE_DTYPE)
    got_st, got_rv = gpu_ctx2.uap_sieve_host(stream, pkts, gs_e, st0)
    assert got_st.tobytes() == want_st.tobytes() and got_rv.tobytes() == want_rv.tobytes()
    cut = len(pkts) // 2
    g_cut = int(np.searchsorted(gs, cut, side="right") - 1)
    cut = int(gs[g_cut])                      # split at a group boundary, then re-feed the later groups' first halves
    st1, rv1 = gpu_ctx2.uap_sieve_host(stream, pkts[:cut], gs[: g_cut + 1], np.zeros(g_cut, dtype=B.SIEVE_DTYPE))
    w1, wr1 = util.sieve_run(orc, "orc", stream, pkts[:cut], gs[: g_cut + 1])
    assert st1.tobytes() == w1.tobytes() and rv1.tobytes() == wr1.tobytes()


def test_oracle_sieve_random_entry_states_vs_reference():
    """Arbitrary piconet states on entry (half-eliminated candidate lists, packet counters next to
    the 1000-packet limit, AFH flags, a UAP that is already known): the restatement must track the
    reference's btbb_uap_from_header from any of them."""
    if not util.have_ref():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    O, R = util.oracle(), util.ref()
    stream, pkts, gs, laps, truth = util.sieve_case(n_slots=500, n_laps=5, ber=0.004, seed=41, mix=("ID", "HV1", "DM1"))
    rng = np.random.default_rng(2718)
    for trial in range(40):
        st = np.zeros(len(gs) - 1, dtype=B.SIEVE_DTYPE)
        for g in range(len(st)):
            flags = 0
            for bit, p in ((2, 0.15), (4, 0.15), (5, 0.1), (9, 0.0), (10, 0.6), (11, 0.3), (12, 0.3)):
                if rng.random() < p:
                    flags |= 1 << bit
            st[g]["flags"] = flags
            st[g]["first_pkt_time"] = rng.integers(0, 1 << 20)
            st[g]["clk_offset"] = rng.integers(0, 64)
            st[g]["packets_observed"] = rng.choice([0, 1, 7, 998, 999, 1000])
            st[g]["total_packets_observed"] = rng.integers(0, 3000)
            st[g]["uap"] = rng.integers(0, 256)
            cand = rng.integers(0, 256, 64).astype(np.int16)
            cand[rng.random(64) < 0.5] = -1
            if rng.random() < 0.5:                      # make the true candidate plausible now and then
                uap, clk0 = truth[laps[g]]
                cand[(clk0 + int(st[g]["first_pkt_time"])) & 63] = uap
            st[g]["clock6_candidates"] = cand
            st[g]["afh_map"] = rng.integers(0, 256, 10)
            st[g]["used_channels"] = int(np.unpackbits(st[g]["afh_map"]).sum())
        a_st, a_rv = util.sieve_run(O, "orc", stream, pkts, gs, states=st)
        b_st, b_rv = util.sieve_run(R, "ref", stream, pkts, gs, states=st)
        assert a_rv.tobytes() == b_rv.tobytes(), trial
        assert a_st.tobytes() == b_st.tobytes(), trial
