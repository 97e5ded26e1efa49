"""Protocol model of the three-CTA cluster kernel (tools/lat2_protocol_model.py): random latency trials must finish without
deadlock, phase aliasing or buffer hazard, with the history margin the launch condition (L >= 12) promises; and every mutation
that removes one handshake must be SEEN by the model (otherwise a green run of it would mean nothing)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import lat2_protocol_model as m  # noqa: E402


@pytest.mark.parametrize("seed", range(16))
def test_random_trials_are_clean(seed):
    sim = m.trial(seed)
    assert all(p.done for p in sim.procs)
    # the prep CTA never reads a history tile the chain CTA wrote fewer than 3 whole steps earlier (L >= 12)
    assert sim.min_history_margin is None or sim.min_history_margin >= 3, sim.min_history_margin


@pytest.mark.parametrize("slow", [{"chain.w0": 10.0}, {"tail.w3": 10.0}, {"prep.w6": 10.0}, {"tail.producer": 10.0}, {"chain.producer": 10.0},
                                  {"prep.producer": 10.0}, {"chain.w5": 0.3, "prep.w1": 0.3}])
def test_a_slow_role_only_slows_the_others(slow):
    sim = m.trial(3, L=12, T=3, slow=slow, max_dil=4)
    assert all(p.done for p in sim.procs)


def _hits(n, **kw):
    hits, last = 0, None
    for s in range(n):
        try:
            m.trial(s, **kw)
        except m.Hazard as e:
            hits, last = hits + 1, str(e)
    return hits, last


def test_mutation_chain_does_not_wait_for_the_h_buffer():
    hits, msg = _hits(6, bug="no_hfree", L=12, T=2, slow={"tail.w3": 10.0})
    assert hits >= 1 and "overwritten" in msg and msg.startswith("h"), msg


def test_mutation_prep_does_not_wait_for_the_tile_buffer():
    hits, msg = _hits(4, bug="no_apfree_wait", L=12, T=2)
    assert hits == 4 and msg.startswith("ap") and "overwritten" in msg, msg


def test_mutation_chain_never_returns_the_tile_buffer():
    hits, msg = _hits(3, bug="no_apfree", L=12, T=2)
    assert hits == 3 and msg.startswith("deadlock"), msg


def test_short_stacks_would_read_history_too_early():
    """Why wn_launch_lat keeps stacks shorter than 12 layers on the single-CTA kernel: at L = 5 the prep CTA runs ahead of the
    history it needs; at L = 8 it is legal with no step to spare."""
    hits, msg = _hits(4, L=5, T=3, max_dil=1, slow={"chain.w3": 3.0})
    assert hits == 4 and "before the chain CTA" in msg, msg
    sim = m.trial(0, L=8, T=3, max_dil=1, slow={"chain.w3": 3.0})
    assert sim.min_history_margin is not None and sim.min_history_margin <= 1
