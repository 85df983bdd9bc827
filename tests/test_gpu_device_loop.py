"""The outer loop that runs ahead of the host (schwz_b200_ras_run): decisions on the device, stop
words, chunked polling, the push at the tail of the iteration, the WHILE-graph CG.  Everything is
checked against the stage-by-stage path (the calls SolverRAS makes, one host synchronisation per
iteration) and the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu(sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")


def _subs(sz, setup, P, **kw):
    ctxs = [sz.Context(0) for _ in range(P)]
    subs = [sz.Ras(ctxs[r], setup, r, **kw) for r in range(P)]
    sz.connect_local(subs, setup)
    return ctxs, subs


def _close(ctxs, subs):
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def _manual_step(subs, it, P):
    for s in subs:
        s.exchange_push(it)
    for s in subs:
        nin, _ = s.neighbors()
        for p in nin:
            s.wait_push_of(subs[int(p)])
        s.exchange_unpack(it)
    for s in subs:
        s.update_boundary()
        s.local_residual()
    norms = [s.residual_norm() for s in subs]
    for s in subs:
        s.local_solve()
        s.restrict()
    return norms


@pytest.mark.parametrize("chunk", [1, 3, 8])
def test_run_ahead_loop_is_bit_identical_to_stage_by_stage(sz, monkeypatch, chunk):
    """same launches, same order per subdomain -> identical doubles, whatever the look-ahead"""
    monkeypatch.setenv("SCHWZ_B200_OUTER_CHUNK", str(chunk))
    n, P, K = 48, 4, 11
    setup = sz.Setup(("laplacian2d", n), P)
    ca, a = _subs(sz, setup, P, local_tol=1e-12)
    cb, b = _subs(sz, setup, P, local_tol=1e-12)
    out = sz.ras_run(a, P, K, tolerance=1e-6, enable_global_check=True, history=True)
    assert out["iters"] == K and not out["converged"]
    hist = np.array([_manual_step(b, it, P) for it in range(K)])
    assert np.array_equal(out["history"], hist)
    sz.refresh_halo(a, P)
    sz.refresh_halo(b, P)
    for r in range(P):
        assert np.array_equal(a[r].x(), b[r].x())
    _close(ca, a)
    _close(cb, b)


def test_converged_run_reports_the_reference_iteration_count(sz, orc):
    """BASELINE.json configs[0] in small: the device-side decision breaks at the same outer
    iteration as the oracle; later launches are no-ops; a second call returns at once"""
    n, P = 32, 2
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs, subs = _subs(sz, setup, P, local_tol=1e-12)
    ob = orc.Problem(*orc.laplacian2d(n), P)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=500, enable_global_check=True)
    ob.run()
    out = sz.ras_run(subs, P, 500, tolerance=1e-6, enable_global_check=True, history=True)
    # metadata.iter_count at the break (the oracle's step counter has moved one past it)
    assert out["converged"] and out["iters"] == int(ob.status(0)["finished_iter"])
    assert out["iters"] == ob.iter_count() - 1
    res, _ = ob.history(0)
    # (the norms inherit the absolute error of the iterates: relative to the first residual)
    np.testing.assert_allclose(out["history"][:, 0], res[:len(out["history"])], rtol=0,
                               atol=1e-9 * res[0])
    x_before = [s.x() for s in subs]
    again = sz.ras_run(subs, P, 500, tolerance=1e-6, enable_global_check=True)
    assert again["converged"] and again["iters"] == 0
    for s, xb in zip(subs, x_before):
        assert np.array_equal(s.x(), xb)
    # after a reset it runs again, to the same count
    for s in subs:
        s.reset()
    third = sz.ras_run(subs, P, 500, tolerance=1e-6, enable_global_check=True)
    assert third["converged"] and third["iters"] == out["iters"]
    _close(ctxs, subs)


def test_warmup_then_timed_runs_continue_one_sequence(sz):
    """bench.py's pattern: ras_run(W) then ras_run(K) continue the same sequence of iterates as
    one ras_run(W + K) (the push left pending by the first call is rewritten, not duplicated)"""
    n, P = 64, 4
    setup = sz.Setup(("laplacian2d", n), P)
    ca, a = _subs(sz, setup, P, local_max_iters=7)
    cb, b = _subs(sz, setup, P, local_max_iters=7)
    h1 = sz.ras_run(a, P, 3, tolerance=1e-9, enable_global_check=True, history=True)["history"]
    h2 = sz.ras_run(a, P, 5, tolerance=1e-9, enable_global_check=True, history=True)["history"]
    h = sz.ras_run(b, P, 8, tolerance=1e-9, enable_global_check=True, history=True)["history"]
    assert np.array_equal(np.vstack([h1, h2]), h)
    # bench.py times the halo kernels between its runs (launches that rewrite / reread the last
    # epoch): the sequence goes on as if nothing had happened
    for kind in (4, 5, 6):
        for s in a:
            s.kernel_time_ms(kind, 3)
    h3 = sz.ras_run(a, P, 4, tolerance=1e-9, enable_global_check=True, history=True)["history"]
    h4 = sz.ras_run(b, P, 4, tolerance=1e-9, enable_global_check=True, history=True)["history"]
    assert np.array_equal(h3, h4)
    # ... and so it does on subdomains whose very first exchange was a timing launch
    cc, c = _subs(sz, setup, P, local_max_iters=7)
    for kind in (5, 6, 4):
        for s in c:
            s.kernel_time_ms(kind, 2)
    for s in c:
        s.reset()
    hc = sz.ras_run(c, P, 8, tolerance=1e-9, enable_global_check=True, history=True)["history"]
    assert np.array_equal(hc, h)
    _close(ca, a)
    _close(cb, b)
    _close(cc, c)


@pytest.mark.parametrize("n", [40, 300])
def test_while_graph_cg_equals_unrolled_and_plain(sz, orc, monkeypatch, n):
    """inexact local solves (--local_tol=0.1 --local_max_iters=70, the regime of the reference's
    run_script): the WHILE-graph CG, the unrolled graph and the plain launch sequence produce the
    same doubles; the oracle agrees to rounding.  n = 40: one-CTA solver; n = 300: multi-kernel CG"""
    P, K = 2, 9
    setup = sz.Setup(("laplacian2d", n), P)
    hists = {}
    for mode in ("1", "0", "plain"):
        if mode == "plain":
            monkeypatch.setenv("SCHWZ_B200_NO_CG_GRAPH", "1")
        else:
            monkeypatch.delenv("SCHWZ_B200_NO_CG_GRAPH", raising=False)
            monkeypatch.setenv("SCHWZ_B200_CG_WHILE", mode)
        ctxs, subs = _subs(sz, setup, P, local_tol=0.1, local_max_iters=70)
        hists[mode] = sz.ras_run(subs, P, K, tolerance=1e-12, enable_global_check=True,
                                 history=True)["history"]
        iters = [s.last_local_iters() for s in subs]
        assert all(0 < i < 70 for i in iters), iters     # stopped by the tolerance, not the cap
        _close(ctxs, subs)
    assert np.array_equal(hists["1"], hists["0"]) and np.array_equal(hists["1"], hists["plain"])
    ob = orc.Problem(*orc.laplacian2d(n), P)
    ob.configure(tolerance=1e-12, local_tol=0.1, local_max_iters=70, max_iters=100,
                 enable_global_check=True)
    for _ in range(K):
        ob.step()
    for r in range(P):
        ref = ob.history(r)[0][:K]
        np.testing.assert_allclose(hists["1"][:, r], ref, rtol=0, atol=1e-8 * ref[0])


def test_exact_local_solves_run_as_one_while_graph(sz, orc, monkeypatch):
    """local_max_iters = -1 (to local_tol): too long to unroll, the WHILE graph needs no host
    polling; iterates agree with the oracle within 1e-10"""
    monkeypatch.delenv("SCHWZ_B200_CG_WHILE", raising=False)
    n, P, K = 200, 2, 6
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs, subs = _subs(sz, setup, P, local_tol=1e-12)
    out = sz.ras_run(subs, P, K, tolerance=1e-6, enable_global_check=True, history=True)
    ob = orc.Problem(*orc.laplacian2d(n), P)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=100, enable_global_check=True)
    for _ in range(K):
        ob.step()
    for r in range(P):
        ref = ob.history(r)[0][:K]
        np.testing.assert_allclose(out["history"][:, r], ref, rtol=0, atol=1e-10 * ref[0])
        l2g = setup.l2g(r)
        xo = ob.x(r)[l2g[:subs[r].local_size]]
        xg = subs[r].x()[:subs[r].local_size]
        assert np.linalg.norm(xg - xo) <= 1e-10 * np.linalg.norm(xo)
    _close(ctxs, subs)


def test_halo_wait_timeout_is_reported(sz, monkeypatch):
    """a neighbour that never publishes its epoch: the bounded wait raises the error word and
    the next host read throws instead of going on with stale halos"""
    monkeypatch.setenv("SCHWZ_B200_HALO_TIMEOUT_MS", "150")
    n, P = 16, 2
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs, subs = _subs(sz, setup, P)
    # subdomain 1 pushes, subdomain 0 does not: 1 waits on 0's flag in vain
    subs[1].exchange_push(0)
    subs[1].exchange_unpack(0, wait_flags=True)
    subs[1].update_boundary()
    subs[1].local_residual()
    with pytest.raises(sz.SchwzError, match="halo exchange timed out"):
        subs[1].residual_norm()
    _close(ctxs, subs)


def test_onesided_run_after_reset_starts_from_clean_buffers(sz, orc):
    """one-sided Put, decentralised flags: after reset() a second run must not scatter the
    converged halo values of the first one (single receive buffer, cleared on entry)"""
    n, P = 32, 4
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs, subs = _subs(sz, setup, P, local_tol=1e-12)
    kw = dict(tolerance=1e-6, enable_onesided=True, conv_decentralized=True,
              enable_global_check=False)
    a = sz.ras_run(subs, P, 3000, **kw)
    assert a["converged"]
    for s in subs:
        s.reset()
    # first exchange of a fresh run with nobody having pushed yet: the buffers must read zero
    subs[0].set_onesided(True)
    subs[0].exchange_unpack(1)
    ls = subs[0].local_size
    assert not subs[0].x()[ls:].any()
    b = sz.ras_run(subs, P, 3000, **kw)
    assert b["converged"]
    sz.refresh_halo(subs, P)
    rsq = sum(s.true_residual_sq() for s in subs)
    assert np.sqrt(rsq) / n <= 1e-4
    # with clean buffers the second run cannot be (much) shorter than the first
    assert b["iters"] >= 0.5 * a["iters"]
    _close(ctxs, subs)
