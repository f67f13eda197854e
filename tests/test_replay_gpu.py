"""
GPU parity of the replay buffer (K5 gather / n-step return, K6 append) through
the C ABI, against (a) the reference's own outputs (tests/golden/replay_*.npz),
(b) the CPU oracle on larger seeded streams, (c) the oracle's restatement of the
on-device index draw (bit-exact replay indices).
"""
import numpy as np
import pytest
import torch

import golden_io
from oracle import replay_oracle as ro

pytestmark = pytest.mark.gpu


def make(case, **kw):
    from rlmd_b200.replay_torch import ReplayBufferTorch
    return ReplayBufferTorch(golden_io.replay_inputs_dict(case), **kw)


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def check_sample(got, want, exact_reward):
    s, a, r, s2, d, eff = [np.asarray(t.cpu().numpy() if torch.is_tensor(t) else t) for t in got]
    ws, wa, wr, ws2, wd, weff = want
    assert np.array_equal(np.atleast_1d(eff).astype(np.int64), np.atleast_1d(weff).astype(np.int64))
    assert np.array_equal(bits(s), bits(ws))
    assert np.array_equal(bits(a), bits(wa))
    assert np.array_equal(bits(s2), bits(ws2))
    assert np.array_equal(d, wd)
    if exact_reward:
        assert np.array_equal(bits(r), bits(wr))
    else:   # torch.sum's summation order over <= n-1 fp32 terms is its own
        np.testing.assert_allclose(r, wr, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("batched_store", [False, True], ids=["store_exp", "store_batch"])
@pytest.mark.parametrize("case", golden_io.REPLAY_CASES, ids=lambda c: c["name"])
def test_follows_reference_fixture(case, batched_store):
    gold = golden_io.load("replay_" + case["name"])
    st = golden_io.replay_stream(case)
    buf = make(case)
    dev = buf.device
    edges = [0] + list(case["events"])
    if edges[-1] != case["fill"]:
        edges.append(case["fill"])
    ev = 0
    for lo, hi in zip(edges[:-1], edges[1:]):
        if batched_store and hi > lo:
            sl = slice(lo, hi)
            buf.store_batch(torch.as_tensor(st["state"][sl], device=dev), torch.as_tensor(st["action"][sl], device=dev),
                            torch.as_tensor(st["reward"][sl], device=dev),
                            torch.as_tensor(st["next_state"][sl], device=dev),
                            torch.as_tensor(st["done"][sl], device=dev))
        else:
            for i in range(lo, hi):
                buf.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i],
                              bool(st["done"][i]))
        while ev < len(case["events"]) and case["events"][ev] == hi:
            got = buf.sample_exp(batch=st["batches"][ev])
            want = [gold[f"ev{ev}_{k}"] for k in ("states", "actions", "rewards", "next_states", "dones", "eff")]
            check_sample(got, want, exact_reward=(case["dyna"] != "A" or case["n"] <= 4))
            ev += 1
    assert ev == len(case["events"])
    assert buf.mem_idx == int(gold["mem_idx"])
    assert np.array_equal(bits(buf.reward_memory[: case["fill"]].cpu().numpy()), bits(gold["reward_memory"]))
    hdr = buf.header.cpu().numpy()
    ends = np.flatnonzero(st["done"])
    assert hdr[0] == case["fill"] and hdr[1] == len(ends)
    assert hdr[2] == ends[0] and hdr[3] == ends[-1] and hdr[4] == ends[-1] + 1


@pytest.mark.parametrize("n,dyna,s,a", [(1, "M", 5, 1), (4, "M", 5, 1), (10, "A", 5, 1), (10, "M", 40, 3), (64, "M", 6, 2)])
def test_large_buffer_against_oracle(n, dyna, s, a):
    """Every slot of a 60k buffer sampled once; batched append in uneven pieces (crosses the 1024-wide scan tiles)."""
    case = dict(name="big", mem=60_000, s=s, a=a, n=n, gamma=0.99, dyna=dyna, batch=512, r0=None,
                lens=[5, 17, 60, 1, 9, 2500, 33, 4], fill=57_777, events=[], seed=100 + n)
    st = golden_io.replay_stream(case)
    buf = make(case)
    ref = ro.ReplayOracle(golden_io.replay_inputs_dict(case))
    dev = buf.device
    cuts = [0, 1, 700, 1024, 1025, 5000, 30_001, case["fill"]]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        sl = slice(lo, hi)
        buf.store_batch(*(torch.as_tensor(st[k][sl], device=dev) for k in ("state", "action", "reward", "next_state", "done")))
    for i in range(case["fill"]):
        ref.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
    ends = np.flatnonzero(st["done"])
    want_start = np.zeros(case["fill"], dtype=np.int64)
    for e0, e1 in zip(ends[:-1], ends[1:]):
        want_start[e0 + 1:e1 + 1] = e0 + 1
    want_start[ends[-1] + 1:] = ends[-1] + 1
    assert np.array_equal(buf.episode_start[: case["fill"]].cpu().numpy(), want_start)
    order = np.random.RandomState(3).permutation(case["fill"])
    got = buf.sample_exp(batch=order)
    want = ref.sample_exp(order)
    check_sample(got, want, exact_reward=True)


def test_ring_overwrite_single_step():
    """multi_steps == 1 wraps like the reference (slot = mem_idx % mem_size, :183); multi-step refuses to."""
    case = dict(golden_io.replay_case("n1"), mem=100, fill=250)
    case["lens"] = [7]
    st = golden_io.replay_stream(dict(case, events=[]))
    buf = make(case)
    ref = ro.ReplayOracle(golden_io.replay_inputs_dict(case))
    for i in range(250):
        args = (st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
        buf.store_exp(*args)
        ref.store_exp(*args)
    assert buf.mem_idx == 250
    batch = np.arange(100)
    check_sample(buf.sample_exp(batch=batch), ref.sample_exp(batch), exact_reward=True)
    multi = make(dict(case, n=3))
    with pytest.raises(RuntimeError):
        for i in range(101):
            multi.store_exp(st["state"][i], st["action"][i], 1.0, st["next_state"][i], False)


@pytest.mark.parametrize("filled,batch,k", [(1000, 256, 3), (300, 256, 2), (64, 64, 1), (1_000_000, 512, 4), (20_000, 8192, 1)])
def test_device_draw_is_the_oracles(filled, batch, k):
    case = dict(name="draw", mem=max(filled, 10), s=2, a=1, n=1, gamma=0.99, dyna="M", batch=batch, r0=None,
                lens=[10], fill=filled, events=[], seed=5)
    buf = make(case, seed=0x1234_5678_9ABC_DEF0)
    dev = buf.device
    z = torch.zeros((filled, 2), dtype=torch.float32, device=dev)
    buf.store_batch(z, z[:, :1], torch.arange(filled, dtype=torch.float32, device=dev), z,
                    torch.zeros(filled, dtype=torch.bool, device=dev))
    idx, s, a, r, s2, d, e = buf.sample_many(k)
    idx = idx.cpu().numpy()
    for b in range(k):
        want = ro.draw_unique(buf.seed, 1, b, filled, batch)
        assert np.array_equal(idx[b], want)
        assert len(set(idx[b].tolist())) == batch and idx[b].min() >= 0 and idx[b].max() < filled
    assert np.array_equal(r.cpu().numpy().astype(np.int64), idx)      # reward memory holds the slot number
    idx2 = buf.sample_many(k)[0].cpu().numpy()
    assert not np.array_equal(idx, idx2), "consecutive draws must differ"
    assert np.array_equal(idx2[0], ro.draw_unique(buf.seed, 2, 0, filled, batch))


def test_device_draw_is_uniform():
    """Every slot equally likely: chi-square of 2000 x 256 draws over 1024 slots."""
    filled, batch, k = 1024, 256, 2000
    case = dict(name="u", mem=filled, s=1, a=1, n=1, gamma=0.9, dyna="M", batch=batch, r0=None, lens=[10], fill=filled,
                events=[], seed=5)
    buf = make(case, seed=99)
    z = torch.zeros((filled, 1), dtype=torch.float32, device=buf.device)
    buf.store_batch(z, z, z[:, 0], z, torch.zeros(filled, dtype=torch.bool, device=buf.device))
    idx = buf.sample_many(k)[0].cpu().numpy()
    counts = np.bincount(idx.ravel(), minlength=filled)
    expect = k * batch / filled
    chi2 = ((counts - expect) ** 2 / expect).sum() / (1 - batch / filled)   # without replacement: variance shrinks
    assert abs(chi2 - (filled - 1)) < 6 * np.sqrt(2 * (filled - 1)), chi2


def test_sample_exp_contract():
    """Return types / shapes of the reference (:390-412); young buffer; error behaviour."""
    case = golden_io.replay_case("n5_M")
    st = golden_io.replay_stream(case)
    buf = make(case)
    for i in range(20):
        buf.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
    with pytest.raises(IndexError):
        buf.sample_exp()                      # fewer than mini_batch_size transitions, multi-step
    for i in range(20, 200):
        buf.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
    s, a, r, s2, d, eff = buf.sample_exp()
    b = case["batch"]
    assert s.shape == (b, 5) and a.shape == (b, 1) and r.shape == (b,) and s2.shape == (b, 5)
    assert d.dtype == torch.bool and eff.dtype == torch.int64 and eff.shape == (b,)
    assert int(eff.min()) >= 1 and int(eff.max()) <= 5
    want = ro.ReplayOracle(golden_io.replay_inputs_dict(case))
    for i in range(200):
        want.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
    check_sample((s, a, r, s2, d, eff), want.sample_exp(buf.last_batch.cpu().numpy()), exact_reward=True)

    one = make(golden_io.replay_case("n1"))
    for i in range(10):
        one.store_exp(st["state"][i], st["action"][i], 1.0, st["next_state"][i], False)
    s, a, r, s2, d, eff = one.sample_exp()     # young buffer: the batch is as long as the buffer (:382-383)
    assert s.shape == (10, 5) and eff.dim() == 0 and int(eff) == 1
    assert sorted(one.last_batch.cpu().tolist()) == list(range(10))


def test_graph_captured_sampler_equals_indexed_sampling():
    """capture_sampler(): every replay draws fresh distinct slots on the device (the oracle's draw for the
    counter value), sees transitions stored after the capture, and returns exactly what sample_exp(batch=those
    slots) returns - the multi-step targets included."""
    case = golden_io.replay_case("n5_M")
    st = golden_io.replay_stream(case)
    buf = make(case, seed=77)
    want = ro.ReplayOracle(golden_io.replay_inputs_dict(case))
    for i in range(300):
        args = (st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
        buf.store_exp(*args)
        want.store_exp(*args)
    run = buf.capture_sampler()
    seen = []
    for rep in range(4):
        if rep == 2:                       # the buffer grows between replays
            for i in range(300, 450):
                args = (st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i], bool(st["done"][i]))
                buf.store_exp(*args)
                want.store_exp(*args)
        out = run()
        idx = buf.last_batch.cpu().numpy().copy()
        filled = 300 if rep < 2 else 450
        assert len(set(idx.tolist())) == case["batch"] and idx.min() >= 0 and idx.max() < filled
        # the warm-up launch consumed counter value c0 (capturing launches nothing): replay r uses c0 + 1 + r
        assert np.array_equal(idx, ro.draw_unique(buf.seed, (1 << 40) + 1 + rep, 0, filled, case["batch"]))
        check_sample(tuple(t.clone() for t in out), want.sample_exp(idx), exact_reward=True)
        seen.append(idx)
    assert not np.array_equal(seen[0], seen[1])
    # k mini-batches per replay
    run3 = buf.capture_sampler(k=3)
    s, a, r, s2, d, eff = run3()
    assert s.shape == (3 * case["batch"], 5) and eff.shape == (3 * case["batch"],)
    idx = buf.last_batch.cpu().numpy().reshape(3, -1)
    for b in range(3):
        assert len(set(idx[b].tolist())) == case["batch"]


@pytest.mark.parametrize("s_dim,n,dyna", [(1, 1, "M"), (5, 1, "M"), (5, 10, "M"), (5, 5, "A"), (6, 3, "M"), (12, 10, "A"),
                                          (13, 2, "M")])
def test_bulk_gather_equals_the_per_call_gather(s_dim, n, dyna):
    """
    >= 4096 samples per launch take the warp-cooperative gather (aligned 16-byte windows through shared
    memory); fewer take the per-thread one.  Same slots -> the same bits, including slots beyond the
    filled part (zeros), the first episode, episode ends and a tail group of fewer than 32 samples.
    """
    mem, fill, b = 40_000, 30_011, 256
    case = dict(name="bulk", mem=mem, s=s_dim, a=1, n=n, gamma=0.97, dyna=dyna, batch=b, r0=None, lens=[10], fill=fill,
                events=[], seed=5)
    buf = make(case, seed=77)
    dev = buf.device
    rs = np.random.RandomState(s_dim * 100 + n)
    ends = np.cumsum(rs.randint(1, 40, size=fill // 10))
    done = np.zeros(fill, dtype=bool)
    done[ends[ends < fill - 5] - 1] = True
    st = torch.as_tensor(rs.standard_normal((fill, s_dim)), device=dev)
    buf.store_batch(st, st[:, :1] * 2, torch.as_tensor(1 + 0.05 * rs.standard_normal(fill), device=dev), st + 1,
                    torch.as_tensor(done, device=dev))
    k = 17
    slots = rs.randint(0, mem, size=(k, b)).astype(np.int64)            # a quarter lies beyond the filled part
    slots[0, :40] = np.arange(40)                                       # the first episode
    slots[1, :64] = np.concatenate([ends[:32] - 1, ends[:32]])          # episode ends and starts
    slots[2, :8] = fill - 1 - np.arange(8)                              # the unfinished run at the end
    slots_t = torch.as_tensor(slots, device=dev)
    big = buf.sample_many(k, batches=slots_t)                           # 4352 samples: bulk
    for j in range(k):
        small = buf.sample_many(1, batches=slots_t[j:j + 1].contiguous())
        for name, g_, w_ in zip("i s a r s2 d e".split(), big, small):
            assert torch.equal(g_[j], w_[0]), (name, j)
    # a tail group: 4100 samples = 128 full warps + 4 samples
    flat = slots_t.reshape(-1)[:4100].contiguous()
    i2, s2_, a2, r2, n2, d2, e2 = buf._sample(1, 4100, flat)
    assert torch.equal(s2_, big[1].reshape(-1, s_dim)[:4100]) and torch.equal(r2, big[3].reshape(-1)[:4100])
    assert torch.equal(n2, big[4].reshape(-1, s_dim)[:4100]) and torch.equal(e2, big[6].reshape(-1)[:4100])
    # drawn on the device (one fused launch at this size) == the same slots supplied
    drawn = buf.sample_many(k)
    assert np.array_equal(drawn[0][3].cpu().numpy(), ro.draw_unique(buf.seed, buf._draws, 3, fill, b))
    again = buf.sample_many(k, batches=drawn[0])
    for name, g_, w_ in zip("i s a r s2 d e".split(), drawn, again):
        assert torch.equal(g_, w_), name


def test_bulk_append_bookkeeping_over_several_blocks():
    """
    The header update of a large append is a multi-block scan of the done flags (partial results meet in the
    spare header words): two appends of > 65536 transitions, the second from an UNALIGNED uint8 flag array,
    against NumPy; the scratch words are zero again afterwards.
    """
    mem = 700_000
    case = dict(name="commit", mem=mem, s=2, a=1, n=3, gamma=0.99, dyna="M", batch=64, r0=None, lens=[10], fill=mem,
                events=[], seed=1)
    buf = make(case)
    dev = buf.device
    rs = np.random.RandomState(11)
    counts = [300_000, 333_337]
    done = rs.rand(sum(counts) + 3) < 0.01
    done[:170_000] = False                          # the first flag sits in the third block's part
    flags = torch.as_tensor(done.astype(np.uint8), device=dev)
    pos = 0
    for i, c in enumerate(counts):
        z = torch.zeros((c, 2), dtype=torch.float32, device=dev)
        off = 0 if i == 0 else 3                    # a view that starts 3 bytes into the allocation
        dn = flags[off + pos: off + pos + c]
        buf.store_batch(z, z[:, :1], torch.ones(c, dtype=torch.float32, device=dev), z, dn)
        pos += c
    want = np.concatenate([done[:counts[0]], done[3 + counts[0]: 3 + counts[0] + counts[1]]])
    ends = np.flatnonzero(want)
    hdr = buf.header.cpu().numpy()
    assert hdr[0] == pos and hdr[1] == len(ends) and hdr[2] == ends[0] and hdr[3] == ends[-1] and hdr[4] == ends[-1] + 1
    assert not hdr[5:].any()
    assert np.array_equal(buf.terminal_memory[:pos].cpu().numpy().astype(bool), want)
    start = np.zeros(pos, dtype=np.int64)
    for e0, e1 in zip(ends[:-1], ends[1:]):
        start[e0 + 1:e1 + 1] = e0 + 1
    start[ends[-1] + 1:] = ends[-1] + 1
    assert np.array_equal(buf.episode_start[:pos].cpu().numpy(), start)
