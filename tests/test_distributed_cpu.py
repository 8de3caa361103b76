"""World-size-2 gloo test (CPU) of the sharding plumbing: ranges, halos, slices and the variable-length gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_windows, S, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypad_b200 import distributed as hd

    ranges = hd.shard_ranges(n_windows, world)
    first, count = ranges[rank]
    t0, tc = hd.timestep_range(first, count, n_windows, S, rank == world - 1)
    # every rank contributes "its" timesteps / windows of two known global arrays
    kmax_global = torch.arange(n_windows + S - 1, dtype=torch.float64) * 0.5
    rec_global = torch.arange(n_windows, dtype=torch.float32) + 7
    tcounts = [hd.timestep_range(f, c, n_windows, S, r == world - 1)[1] for r, (f, c) in enumerate(ranges)]
    kmax = hd.gather_concat(kmax_global[t0:t0 + tc].clone(), tcounts)
    rec = hd.gather_concat(rec_global[first:first + count].clone(), [c for _, c in ranges])
    ok = torch.equal(kmax, kmax_global) and torch.equal(rec, rec_global)

    class FakeScorer:
        pass

    fs = FakeScorer()
    fs.S = S
    sh = hd.ShardedScorer(fs)
    f2, c2, h0, lo, hi = sh.plan(n_windows)
    # long arrays are dealt out in whole blocks of 1024 positions (the sharded find_anomalies), short ones balanced to the window
    want = hd.shard_ranges_aligned(n_windows, world) if n_windows >= world * 8 * hd.BLOCK else ranges
    ok = ok and (f2, c2) == want[rank] and h0 == max(0, f2 - (S - 1)) and lo == h0 and hi - lo - S == f2 + c2 - h0
    al = hd.shard_ranges_aligned(n_windows, world)
    ok = ok and sum(c for _, c in al) == n_windows and all(f % hd.BLOCK == 0 or c == 0 for f, c in al)
    ok = ok and all(al[i][0] + al[i][1] == al[i + 1][0] for i in range(world - 1))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows,S", [(1001, 100), (64, 100), (999900, 100), (7, 3)])
def test_shard_gather_world2(n_windows, S):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_windows, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_ranges_cover_exactly():
    from hypad_b200 import distributed as hd

    for n in (1, 2, 7, 100, 8540, 999900):
        for world in (1, 2, 3, 4, 8):
            r = hd.shard_ranges(n, world)
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in r) - min(c for _, c in r) <= 1
            total_t = sum(hd.timestep_range(f, c, n, 100, i == world - 1)[1] for i, (f, c) in enumerate(r))
            assert total_t == n + 99


# ---- bulk sweep: sharding by signal (BASELINE config 5) -----------------------------------------------------
def bundled_length_multiset():
    """A length multiset shaped like the bundled data (80 NASA-test, 46 NAB, 367 YAHOO signals), seeded."""
    rng = np.random.default_rng(5)
    return np.concatenate([rng.integers(2000, 8700, 80), rng.integers(1100, 22700, 46), rng.integers(1420, 1700, 367)]).tolist()


def test_assign_signals_is_a_balanced_partition():
    from hypad_b200.sweep import assign_signals

    lengths = bundled_length_multiset()
    n = [t - 100 for t in lengths]
    for world in (1, 2, 4, 8):
        plan = assign_signals(n, world)
        assert sorted(i for ids in plan for i in ids) == list(range(len(n)))  # every signal exactly once
        loads = [sum(n[i] for i in ids) for ids in plan]
        assert max(loads) - min(loads) <= max(n)  # LPT: no rank is ahead by more than one signal
        assert plan == assign_signals(n, world)  # deterministic
        for ids in plan:
            assert [n[i] for i in ids] == sorted((n[i] for i in ids), reverse=True)  # longest first on every rank
    assert assign_signals([], 3) == [[], [], []]
    assert assign_signals([7, 7, 7], 2) == [[0, 2], [1]]  # ties: lower id first, lower rank first
    with pytest.raises(ValueError):
        assign_signals([1], 0)


def _sweep_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypad_b200.sweep import SignalSweep

    lengths = bundled_length_multiset()[:40] + [100, 37]  # two signals too short for a window
    signals = [np.zeros(t) for t in lengths]
    indices = [np.arange(t) for t in lengths]

    class Sweep(SignalSweep):  # the device work replaced by a recognisable stand-in: the plumbing is what runs here
        def score_local(self, signals, indices, ids, combination="uncertainty", rec_error_type="dtw", keep_scores=False):
            self.scored = list(ids)
            return {i: {"intervals": np.array([[i, len(signals[i]), self.rank]], dtype=np.float64)} for i in ids}

    sw = Sweep(scorers=None)
    res = sw.run(signals, indices)
    plan = sw.plan(lengths)
    ok = sw.scored == plan[rank] and sorted(res) == list(range(len(lengths)))
    for r, ids in enumerate(plan):
        for i in ids:
            ok = ok and res[i].tolist() == [[i, lengths[i], r]]
    ok = ok and res[40].shape == (0, 3) and res[41].shape == (0, 3) and not any(i in (40, 41) for ids in plan for i in ids)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_signal_sweep_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n,S,rows", [(2500, 123, True), (1400, 100, False), (130, 100, False), (7, 3, True)])
def test_plan_and_gather_layout_round_trip(world, n, S, rows):
    """The plan of every rank of a run (owned windows, halo, resident samples / rows, owned positions of the kmax array) and the
    layout of the padded all-gathers that put the ranks' slices back together (ShardedScorer.gather_full), on CPU tensors with
    a stand-in for the exchange: slices cut the way the ranks hold them must come back as the global arrays."""
    from hypad_b200 import distributed as hd

    class FakeScorer:
        pass

    fs = FakeScorer()
    fs.S = S
    kmax_g = torch.arange(n + S - 1, dtype=torch.float64) * 0.25 + 1
    rec_g = torch.arange(n, dtype=torch.float32) + 1000
    fin_g = torch.arange(n, dtype=torch.float64) + 5000
    ranges = hd.shard_ranges(n, world)
    slots = {}

    class Comm:
        def __init__(self, rank):
            self.rank, self.world = rank, world

        def all_gather(self, buf):  # every rank's buffer for this call was deposited beforehand
            return torch.stack(slots[(buf.dtype, buf.numel())])

    shs, outs, covered, pos = [], [], 0, 0
    for r in range(world):
        sh = hd.ShardedScorer(fs, rank=r, world=world, comm=Comm(r))
        first, count, h0, lo, hi = sh.plan_rows(n) if rows else sh.plan(n)
        assert (first, count) == ranges[r] and h0 == max(0, first - (S - 1)) and lo == h0
        assert hi == first + count + (0 if rows else S)
        covered += count
        t0, tc = sh.position_ranges(n)[r]
        assert t0 == pos and tc == count + (S - 1 if r == world - 1 else 0)
        pos += tc
        outs.append({"final_local": fin_g[first:first + count], "rec_local": rec_g[first:first + count],
                     "unorm_local": rec_g[first:first + count] * 2, "critic_scores_local": fin_g[first:first + count] * 3,
                     "kmax_local": kmax_g[t0:t0 + tc]})
        shs.append(sh)
    assert covered == n and pos == n + S - 1
    wmax = max(c for _, c in ranges)
    pmax = max(c for _, c in shs[0].position_ranges(n))
    # one array at a time: the stand-in hands back what all ranks would have contributed to this call
    for r in range(world):
        for key, width, want in (("final_local", wmax, fin_g), ("rec_local", wmax, rec_g), ("kmax_local", pmax, kmax_g)):
            slots[(outs[0][key].dtype, width)] = [hd._pad_to(outs[q][key], width) for q in range(world)]
            if key == "kmax_local":
                got = hd._concat_rows(shs[r].comm.all_gather(hd._pad_to(outs[r][key], width)), [c for _, c in shs[r].position_ranges(n)])
            else:
                got = shs[r]._gather_windows(outs[r][key], n)
            assert torch.equal(got, want), (r, key)


def test_halo_from_strips_walks_over_short_neighbours():
    """The smoothing halo is cut from the neighbours' edge strips; a neighbour shorter than the halo hands over all it has and
    the walk goes on to the next rank (hypad_b200.scoring._halo_from_strips)."""
    from hypad_b200.scoring import _halo_from_strips

    full = torch.arange(40, dtype=torch.float64)
    for lens, H in (([10, 10, 10, 10], 4), ([17, 2, 1, 20], 5), ([3, 3, 3, 31], 6), ([40], 3), ([0, 20, 0, 20], 4)):
        ranges, p = [], 0
        for ln in lens:
            ranges.append((p, ln))
            p += ln
        strips = torch.zeros(len(lens), 2, H, dtype=torch.float64)
        for r, (p0, ln) in enumerate(ranges):
            k = min(ln, H)
            if k:
                strips[r, 0, :k] = full[p0:p0 + k]
                strips[r, 1, H - k:] = full[p0 + ln - k:p0 + ln]
        for r, (p0, ln) in enumerate(ranges):
            nl, nr = min(H, p0), min(H, 40 - (p0 + ln))
            left, right = _halo_from_strips(strips, ranges, r, nl, nr)
            got = torch.cat(left + [full[p0:p0 + ln]] + right)
            assert torch.equal(got, full[p0 - nl:p0 + ln + nr]), (lens, r)
