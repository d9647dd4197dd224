"""GPU parity tests: the drop-in library (CUDA path) against the reference, through the public
C API (harness/harness.c).  Where oracle/_ref is present both libraries run in lock-step in
this process; the committed golden vectors (tools/make_golden.py) are always checked too.

Tolerance: 1e-6 relative on node states and chi2 (BASELINE.json north_star); observed
differences are ~1e-9 (different elimination arithmetic order, GPU sin/cos).
"""
import numpy as np
import pytest

from aprilsam_b200 import harness as H
from conftest import golden

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def rel_state_err(a, b):
    """max |a-b| / max(1, |b|) with theta compared modulo 2 pi."""
    d = a - b
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return float(np.max(np.abs(d) / np.maximum(1.0, np.abs(b))))


def have_ref():
    return H.available("reference")


def test_chi2_kernel(m3500):
    g = golden("m3500_batch.npz")
    with H.Harness("b200") as h:
        h.load_full(m3500)
        c0 = h.chi2()
    assert abs(c0 - g["chi2"][0]) <= RTOL * g["chi2"][0]


def test_m3500_batch_six_iterations(m3500):
    g = golden("m3500_batch.npz")
    with H.Harness("b200") as h:
        h.load_full(m3500)
        for it in range(6):
            h.batch()
            assert np.array_equal(h.ordering(), g["ordering"]) or it > 0
            c = h.chi2()
            assert abs(c - g["chi2"][it + 1]) <= RTOL * g["chi2"][it + 1], (it, c, g["chi2"][it + 1])
            err = rel_state_err(h.states(), g["states"][it])
            assert err < RTOL, (it, err)
        assert np.array_equal(h.tree_parents(), g["tree_parents"])


def test_small_graphs_batch(m3500):
    """Ragged / tiny inputs: 1, 2, 3, 7, 50 poses."""
    if not have_ref():
        pytest.skip("reference oracle not built on this box")
    for n in (1, 2, 3, 7, 50):
        sub = m3500.head(n)
        with H.Harness("b200") as a, H.Harness("reference") as b:
            a.load_full(sub)
            b.load_full(sub)
            for _ in range(3):
                a.batch()
                b.batch()
                assert rel_state_err(a.states(), b.states()) < RTOL, n
                cb = b.chi2()
                assert abs(a.chi2() - cb) <= RTOL * max(1.0, cb), n
            assert np.array_equal(a.ordering(), b.ordering())


def test_tutorial_graph():
    g = golden("tutorial.npz")
    d = H.PoseGraphData(g["init"], g["ea"], g["eb"], g["ez"], g["eW"])
    for mode, batch_only in (("inc", False), ("batch", True)):
        with H.Harness("b200") as h:
            h.replay_begin(d)
            chi2, _, _ = h.replay_to(6, batch_only=batch_only)
            assert rel_state_err(h.states(), g[f"{mode}_states"]) < RTOL
            assert abs(chi2[-1] - g[f"{mode}_chi2"][-1]) <= RTOL * max(1.0, g[f"{mode}_chi2"][-1])


@pytest.mark.parametrize("nsteps", [400])
def test_m3500_replay_lockstep(m3500, nsteps):
    """Every step: states of ALL nodes, chi2, naffected and start_over must match."""
    if not have_ref():
        pytest.skip("reference oracle not built on this box")
    with H.Harness("b200") as a, H.Harness("reference") as b:
        a.replay_begin(m3500)
        b.replay_begin(m3500)
        worst = 0.0
        for k in range(1, nsteps + 1):
            ca, _, ia = a.replay_to(k)
            cb, _, ib = b.replay_to(k)
            assert ia[0][0] == ib[0][0], f"step {k}: naffected {ia[0][0]} vs {ib[0][0]}"
            assert ia[0][1] == ib[0][1], f"step {k}: start_over {ia[0][1]} vs {ib[0][1]}"
            err = rel_state_err(a.states(), b.states())
            worst = max(worst, err)
            assert err < RTOL, f"step {k}: state err {err}"
            assert abs(ca[0] - cb[0]) <= RTOL * max(1.0, cb[0]), f"step {k}: chi2 {ca[0]} vs {cb[0]}"
        print("worst relative state error over", nsteps, "steps:", worst)


def test_m3500_replay_full_golden(m3500):
    """Whole 3500-step replay against the committed per-step chi2 / counters and checkpoints."""
    g = golden("m3500_replay.npz")
    with H.Harness("b200") as h:
        h.replay_begin(m3500)
        done = 0
        for cp in g["checkpoints"]:
            chi2, _, info = h.replay_to(int(cp))
            n = len(chi2)
            assert np.array_equal(info[:, 0], g["naffected"][done:done + n]), f"naffected differs before step {cp}"
            assert np.array_equal(info[:, 1], g["start_over"][done:done + n]), f"start_over differs before step {cp}"
            ref = g["chi2"][done:done + n]
            assert np.all(np.abs(chi2 - ref) <= RTOL * np.maximum(1.0, ref)), f"chi2 differs before step {cp}"
            err = rel_state_err(h.states(), g[f"states_{int(cp)}"])
            assert err < RTOL, (int(cp), err)
            done += n
        assert abs(chi2[-1] - 68.965607796) < 1e-6
