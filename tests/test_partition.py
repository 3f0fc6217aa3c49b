"""Host logic of the xi-slab partition (no GPU): the reference's equal-length rule and the cost-balanced cut."""
import numpy as np
import pytest

from qpad_b200.pipeline import balanced_partition, slab_partition, split_beam


def _tiles(parts, nz):
    return parts[0][0] == 0 and sum(n for _, n in parts) == nz and all(a + n == b for (a, n), (b, _) in zip(parts[:-1], parts[1:]))


def test_equal_length_rule_of_the_reference():
    """options_class.f03:103-106: nz / S each, the remainder goes to the first stages"""
    assert slab_partition(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert slab_partition(2048, 8) == [(256 * k, 256) for k in range(8)]
    assert _tiles(slab_partition(438, 4), 438)


@pytest.mark.parametrize("G", [2, 3, 4, 8, 16, 32])
def test_balanced_partition_minimises_the_slowest_slab(G):
    rng = np.random.default_rng(G)
    nz = 2048
    cost = 88.0 + 2.0 * rng.random(nz)
    cost[384:640] += 46.0                      # the wake: two predictor-corrector iterations per slice
    parts = balanced_partition(cost, G, min_len=4)
    assert len(parts) == G and _tiles(parts, nz) and all(n >= 4 for _, n in parts)
    worst = max(cost[a:a + n].sum() for a, n in parts)
    uniform = max(cost[a:a + n].sum() for a, n in slab_partition(nz, G))
    ideal = cost.sum() / G
    assert worst <= uniform + 1e-9
    assert worst <= ideal + cost.max()        # within one slice of the lower bound
    # brute-force optimum for the small cases
    if G <= 3:
        best = np.inf
        cs = np.concatenate([[0.0], np.cumsum(cost)])
        if G == 2:
            best = min(max(cs[c], cs[nz] - cs[c]) for c in range(4, nz - 3))
        else:
            for c1 in range(4, nz - 7, 8):
                c2 = np.arange(c1 + 4, nz - 3)
                best = min(best, np.min(np.maximum(np.maximum(cs[c1], cs[c2] - cs[c1]), cs[nz] - cs[c2])))
        assert worst <= best * (1 + 1e-3) + cost.max() * (G == 3)


def test_balanced_partition_edge_cases():
    assert balanced_partition(np.ones(10), 5) == [(0, 2), (2, 2), (4, 2), (6, 2), (8, 2)]
    assert balanced_partition([5, 1, 1, 1, 1, 1, 1, 1], 4) == [(0, 2), (2, 2), (4, 2), (6, 2)]     # min_len binds
    assert balanced_partition(np.ones(7), 1) == [(0, 7)]
    p = balanced_partition(np.zeros(64), 4)                                                   # no information: still a tiling
    assert _tiles(p, 64) and len(p) == 4
    with pytest.raises(ValueError):
        balanced_partition(np.ones(5), 4)


def test_split_beam_follows_the_partition():
    """a beam particle belongs to the stage whose slab holds its xi (beam/part3d_comm.f03), also for unequal slabs"""
    nz, dxi = 32, 0.25
    parts = [(0, 5), (5, 17), (22, 10)]
    rng = np.random.default_rng(0)
    bx = np.zeros((200, 3)); bx[:, 2] = rng.uniform(0, nz * dxi, 200)
    bp, bq = rng.standard_normal((200, 3)), rng.standard_normal(200)
    pieces = split_beam(bx, bp, bq, nz, dxi, 3, parts=parts)
    assert sum(len(q) for _, _, q in pieces) == 200
    for (noff, n), (x, _, _) in zip(parts, pieces):
        assert np.all(x[:, 2] >= noff * dxi) and np.all(x[:, 2] < (noff + n) * dxi)
    same = split_beam(bx, bp, bq, nz, dxi, 4)
    for (noff, n), (x, _, _) in zip(slab_partition(nz, 4), same):
        assert np.all(x[:, 2] >= noff * dxi) and np.all(x[:, 2] < (noff + n) * dxi)


class _StubSim:
    def __init__(self, trace, sweep_ms):
        self.trace, self.sweep_ms = np.asarray(trace, dtype=float), sweep_ms

    def sweep_profile(self, reset=False):
        return {"ns_total": self.sweep_ms * 1e6 * 4}          # measured_partition divides by its nwaves (4 below)

    def slice_trace(self):
        return self.trace, np.ones(len(self.trace), dtype=np.int32)


class _StubPipeline:
    """the surface of LocalPipeline that pipeline.measured_partition uses"""

    def __init__(self, parts, ns_per_slice, other_ms):
        self.parts, self.G, self.S, self.world, self.dist = parts, len(parts), len(parts), 1, None
        self.sims = [_StubSim(ns_per_slice[a:a + n], ns_per_slice[a:a + n].sum() * 1e-6) for a, n in parts]
        self.other_ms, self._ev_on, self.waves = other_ms, False, 0

    def fill(self): pass
    def wave(self): self.waves += 1
    def sync(self): pass
    def trace_reset(self): pass

    def event_report(self):
        return [{"begun>swept": s.sweep_ms, "tail>pre": o, "pre>got_back": 3.3, "w_fwd>got_fwd": 1.1} for s, o in zip(self.sims, self.other_ms)]


def test_closed_loop_partition_equalises_busy_time():
    """pipeline.measured_partition: slabs are re-cut from what the running pipeline measures -- per-stage busy time (waits excluded) and
    the per-slice cost profile -- until the busy times agree; the non-sweep work of a stage (beam kernels) counts"""
    from qpad_b200.pipeline import measured_partition
    nz = 2048
    ns = np.full(nz, 90e3)                       # 90 us per slice ...
    ns[1100:1500] = 125e3                        # ... 125 us inside the wake
    cfg = {"nz": nz}
    parts = slab_partition(nz, 4)
    other = [4.0, 4.0, 0.1, 0.1]                 # ms of beam work per wave in the stages that hold the beam
    for rnd in range(4):
        lp = _StubPipeline(parts, ns, other)
        new, busy, spread = measured_partition(lp, cfg, nwaves=4, nwarm=1)
        assert lp.waves == 5 and len(busy) == 4
        assert all(abs(b - (s.sweep_ms + o)) < 1e-9 for b, s, o in zip(busy, lp.sims, other))       # the two waits are not busy time
        if new is None:
            break
        assert _tiles(new, nz)
        parts = new
    assert new is None and spread < 0.03 and rnd <= 2
    assert parts != slab_partition(nz, 4) and parts[2][1] < parts[3][1]      # the stage that holds the wake got a shorter slab than the quiet tail
