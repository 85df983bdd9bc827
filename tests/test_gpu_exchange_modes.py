"""The four one-sided data paths of the reference (Settings::comm_settings enable_put /
enable_get x enable_one_by_one, source/restricted_schwarz.cpp:753-851) and the centralised
tree convergence protocol (include/conv_tools.hpp:147-209) on the GPU, through the C ABI.

* data movement: with x[own] set to a known function of the global index, ONE exchange of
  every variant must leave exactly the owners' values in every overlap / halo slot
  (bit-exact: the exchange only moves doubles);
* protocol: an asynchronous run of every variant stops by itself and meets the threshold the
  oracle reaches (the north star's criterion for async mode).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODES = ["put", "get", "put-one-by-one", "get-one-by-one"]


@pytest.fixture(autouse=True)
def _need_gpu(sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")


def _build(sz, setup, P, **kw):
    ctxs = [sz.Context(0) for _ in range(P)]
    subs = [sz.Ras(ctxs[r], setup, r, **kw) for r in range(P)]
    sz.connect_local(subs, setup)
    return ctxs, subs


def _close(ctxs, subs):
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", ["strips", "regular2d", "ani4_metis"])
def test_one_exchange_moves_exactly_the_owners_values(sz, ani4, mode, case):
    if case == "strips":
        P, N = 4, 24 * 24
        setup = sz.Setup(("laplacian2d", 24), P)
    elif case == "regular2d":
        P, N = 4, 16 * 16
        setup = sz.Setup(("laplacian2d", 16), P, part=sz.partition_regular2d(N, P))
    else:
        P, N = 8, len(ani4[0]) - 1
        setup = sz.Setup(ani4, P, part=sz.partition_metis(ani4[0], ani4[1], P))
    ctxs, subs = _build(sz, setup, P)
    f = lambda gid: np.sin(gid.astype(np.float64)) * 1e3 + gid      # noqa: E731
    fr = setup.first_row()
    for r, s in enumerate(subs):
        s.set_exchange_mode(mode)
        s.set_x_own(f(np.arange(fr[r], fr[r + 1])))
    for s in subs:
        s.exchange_push(1)
    for s in subs:
        s.sync()                      # every pack / put has landed
    for s in subs:
        s.exchange_unpack(1)
    for r, s in enumerate(subs):
        l2g = setup.l2g(r)
        got = s.x()
        # every slot whose owner is a neighbour now holds the owner's value; own slots unchanged
        assert np.array_equal(got, f(l2g)), (mode, case, r)
    _close(ctxs, subs)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("conv", ["decentralized", "centralized-tree"])
def test_async_run_stops_and_meets_the_threshold(sz, orc, mode, conv):
    n, P, tol = 24, 4, 1e-6
    mat = orc.laplacian2d(n)
    ob = orc.Problem(*mat, P)
    ob.configure(tolerance=tol, max_iters=4000, enable_onesided=True,
                 remote_comm_type=mode.split("-")[0], enable_one_by_one=mode.endswith("one-by-one"),
                 global_convergence_type=conv)
    ob.run()
    _, fro = ob.final_residual()
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs, subs = _build(sz, setup, P)
    out = sz.ras_run(subs, P, 4000, tolerance=tol, enable_onesided=True,
                     conv_decentralized=(conv == "decentralized"), exchange=mode)
    assert out["converged"] and out["iters"] < 4000
    for s in subs:
        s.set_exchange_mode("put")
        s.exchange_push(0)
    for s in subs:
        s.sync()
    for s in subs:
        s.exchange_unpack(0)
    rel = np.sqrt(sum(s.true_residual_sq() for s in subs)) / np.sqrt(n * n)
    assert rel <= max(10 * fro["relative"], 50 * tol)
    _close(ctxs, subs)


def test_tree_protocol_step_by_step(sz):
    """conv words after each call, P = 5 (ranks 2, 3, 4 are leaves, rank 1 has two children,
    rank 0 has children 1 and 2): nothing is announced until every leaf has pushed up, then
    the root's flag travels down one level per call."""
    P = 5
    setup = sz.Setup(("laplacian2d", 10), P)
    ctxs, subs = _build(sz, setup, P)

    def call(flags):
        out = []
        for s, f in zip(subs, flags):
            s.conv_tree(f)
        for s in subs:
            out.append(s.conv_count())
        return out

    assert call([0, 0, 0, 0, 0]) == [0] * 5
    assert call([1, 1, 0, 1, 1]) == [0] * 5          # leaf 2 missing: root cannot fire
    # now everybody is locally converged: 2 pushes to 0; 1 (children pushed above) pushes to
    # 0 in the same sweep because its conv[0], conv[1] were set by 3 and 4 one call earlier
    got = call([1, 1, 1, 1, 1])
    assert got[0] in (0, P)                            # root fires once both children are in
    for _ in range(4):
        got = call([1, 1, 1, 1, 1])
    assert got == [P] * 5
    _close(ctxs, subs)


def test_accumulate_protocol_step_by_step(sz):
    """--enable_decentralized_accumulate (include/conv_tools.hpp:230-247): every call in which a
    subdomain is locally converged adds 1 to word 0 of EVERY subdomain's flags; the count a
    subdomain reports is its word 0.  Calls are serialised here (sync after each), so the
    counts must equal the sequential model the oracle uses."""
    P = 4
    setup = sz.Setup(("laplacian2d", 12), P)
    ctxs, subs = _build(sz, setup, P)
    model = [0] * P                                    # word 0 of every subdomain
    for flags in ([0, 0, 0, 0], [1, 0, 0, 0], [1, 0, 1, 0], [1, 1, 1, 1], [0, 0, 0, 1]):
        for r, (s, f) in enumerate(zip(subs, flags)):
            s.conv_accumulate(f)
            if f:
                model = [m + 1 for m in model]
            assert s.conv_count() == model[r], (flags, r)   # conv_count synchronises
    assert model == [8] * P                            # the counter steps over P: upstream's quirk
    _close(ctxs, subs)
    # through the loop (option plumbing only: whether the count ever EQUALS P depends on the
    # order in which racing subdomains add and read, upstream as here)
    setup = sz.Setup(("laplacian2d", 20), 2)
    ctxs, subs = _build(sz, setup, 2)
    out = sz.ras_run(subs, 2, 6, tolerance=1e-5, enable_onesided=True, conv_decentralized=True,
                     enable_accumulate=True)
    assert out["iters"] == 6 and not out["converged"]
    _close(ctxs, subs)


@pytest.mark.parametrize("mode", ["put", "get"])
def test_mixed_precision_wire_format_rounds_to_float(sz, mode):
    """use_mixed_precision: gathered exchanges carry floats (half the NVLink bytes); the
    receiver sees exactly (double)(float)value."""
    P, n = 4, 24
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs, subs = _build(sz, setup, P, use_mixed_precision=True)
    f = lambda gid: np.sin(gid.astype(np.float64)) * 1e3 + gid / 7.0      # noqa: E731
    fr = setup.first_row()
    for r, s in enumerate(subs):
        s.set_exchange_mode(mode)
        s.set_x_own(f(np.arange(fr[r], fr[r + 1])))
    for s in subs:
        s.exchange_push(1)
    for s in subs:
        s.sync()
    for s in subs:
        s.exchange_unpack(1)
    for r, s in enumerate(subs):
        l2g = setup.l2g(r)
        want = f(l2g)
        want[s.local_size:] = want[s.local_size:].astype(np.float32).astype(np.float64)
        assert np.array_equal(s.x(), want), (mode, r)
    _close(ctxs, subs)


def test_mixed_precision_sync_run_matches_oracle(sz, orc):
    """Float rounding on the wire makes the run discontinuous in its inputs: a 1-ulp (double)
    difference in a halo value that sits on a float rounding boundary becomes a 6e-8 relative
    jump.  So the GPU run is compared with the oracle at the float level (1e-5 on the residual
    history, 1e-6 on the iterates), not at 1e-10, and the stopping iteration may move by one."""
    n, P = 24, 4
    part = orc.partition_regular2d(n * n, P)
    ob = orc.Problem(*orc.laplacian2d(n), P, part=part)
    ob.configure(tolerance=1e-5, local_tol=1e-12, max_iters=400, enable_global_check=True,
                 use_mixed_precision=True)
    iters = ob.run()
    setup = sz.Setup(("laplacian2d", n), P, part=part)
    ctxs, subs = _build(sz, setup, P, local_tol=1e-12, use_mixed_precision=True)
    out = sz.ras_run(subs, P, 400, tolerance=1e-5, enable_global_check=True, history=True)
    assert out["converged"] and abs(out["iters"] - iters) <= 1, (out["iters"], iters)
    _, gres = ob.history(0)
    k = min(len(gres), len(out["history"]))
    np.testing.assert_allclose(out["history"].sum(axis=1)[:k], gres[:k], rtol=0, atol=1e-5 * gres[0])
    if out["iters"] == iters:
        for r in range(P):
            l2g = setup.l2g(r)
            xo = ob.x(r)[l2g]
            assert np.linalg.norm(subs[r].x() - xo) <= 1e-6 * np.linalg.norm(xo)
    # and it really is the float run, not the fp64 one
    ob64 = orc.Problem(*orc.laplacian2d(n), P, part=part)
    ob64.configure(tolerance=1e-5, local_tol=1e-12, max_iters=400, enable_global_check=True)
    ob64.run()
    g64 = ob64.history(0)[1]
    assert abs(out["history"].sum(axis=1)[5] - gres[5]) < abs(g64[5] - gres[5])
    _close(ctxs, subs)
