"""Discrete-event models of the barrier protocols of the two kernels that were written without a GPU at hand
(csrc/gemm_pair.cuh, csrc/gemm_rows_seeded.cuh).  They model my reading of the mbarrier rules - arrival counts, phase
parity, transaction bytes that may complete before they are expected - not the hardware, and check under many random
interleavings (asynchronous TMA completions and MMA commits arrive late and out of step with the issuing threads) that

  * nobody deadlocks,
  * an MMA never reads a ring slot before BOTH CTAs' bytes of that k-block landed, and no producer overwrites a slot
    that an MMA in flight still reads,
  * an accumulator is never overwritten before BOTH epilogues released it, and every epilogue reads the tile it expects,
  * (seeded kernel) the epilogue's phase A / grid barrier / phase B bookkeeping stays in step with the producer and the
    MMA issuer, which only see one tile sequence.

A wrong arrival count or parity in the kernels' source would show up here as a deadlock or a stale read."""
import random

import pytest


class MBar:
    """mbarrier: `count` arrivals + a transaction-byte counter complete a phase; waiters test the phase parity."""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self, n=1):
        assert self.pending >= n, "more arrivals than the barrier was initialised for"
        self.pending -= n
        self._maybe_flip()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):          # may run before the expect: the counter goes negative meanwhile
        self.tx -= nbytes
        self._maybe_flip()

    def passed(self, parity):               # mbarrier.try_wait.parity
        return self.phase != parity


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.actors, self.events = [], []   # events: [delay, fn] fired in FIFO order per queue once delay hits 0

    def spawn(self, gen):
        self.actors.append([gen, None])     # [generator, condition it is blocked on]

    def later(self, fn, queue):
        queue.append(fn)

    def run(self, queues, limit=2_000_000):
        steps = 0
        while self.actors or any(queues):
            steps += 1
            assert steps < limit, "livelock"
            choices = []
            for a in self.actors:
                if a[1] is None or a[1]():
                    choices.append(("actor", a))
            for q in queues:
                if q:
                    choices.append(("event", q))
            assert choices, "DEADLOCK: " + ", ".join(getattr(a[0], "__name__", "?") for a in self.actors)
            kind, obj = self.rng.choice(choices)
            if kind == "event":
                obj.pop(0)()
            else:
                try:
                    obj[1] = next(obj[0])
                except StopIteration:
                    self.actors.remove(obj)


# ------------------------------------------------------------------------------------------------
# gemm_pair_kernel: two CTAs, one MMA issuer (leader), ring of `stages`, two accumulators
# ------------------------------------------------------------------------------------------------
def run_pair(seed, stages, nkb, ntiles):
    sim = Sim(seed)
    A_B = 32768                                             # bytes one CTA loads per k-block (16 KB A + 16 KB B half)
    full = [MBar(1) for _ in range(stages)]                 # leader only
    empty = [[MBar(1) for _ in range(stages)] for _ in range(2)]
    tfull = [[MBar(1) for _ in range(2)] for _ in range(2)]
    tempty = [MBar(256) for _ in range(2)]                  # leader only: 128 threads of each CTA
    slot = [[None] * stages for _ in range(2)]              # what the ring slot of each CTA holds: (tile, kb)
    reading = [[0] * stages for _ in range(2)]              # MMAs in flight that still read the slot
    accum = [[None, None], [None, None]]                    # per CTA: tile each accumulator holds
    acc_busy = [0, 0]                                       # MMAs in flight writing accumulator a
    tma_q = [[], []]                                        # async completions, FIFO per CTA
    mma_q = []                                              # async MMA completions + commits, in issue order
    seen = [[], []]

    def producer(c):
        stage, phase = 0, 0
        for t in range(ntiles):
            for kb in range(nkb):
                yield lambda s=stage, p=phase: empty[c][s].passed(p ^ 1)
                assert reading[c][stage] == 0, "slot refilled while an MMA still reads it"
                if c == 0:
                    full[stage].arrive_expect_tx(2 * A_B)
                def land(s=stage, t=t, kb=kb):
                    slot[c][s] = (t, kb)
                    full[s].complete_tx(A_B)
                sim.later(land, tma_q[c])
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    def mma():
        stage, phase, acc, acc_phase = 0, 0, 0, 0
        for t in range(ntiles):
            yield lambda a=acc, p=acc_phase: tempty[a].passed(p ^ 1)
            for kb in range(nkb):
                yield lambda s=stage, p=phase: full[s].passed(p)
                assert slot[0][stage] == (t, kb) and slot[1][stage] == (t, kb), "MMA read a slot before both halves landed"
                for c in range(2):
                    reading[c][stage] += 1
                acc_busy[acc] += 1
                def done(s=stage, a=acc):                   # the MMAs of this k-block complete, then their commit arrives
                    for c in range(2):
                        reading[c][s] -= 1
                        empty[c][s].arrive()                # tcgen05.commit.cta_group::2 multicast
                    acc_busy[a] -= 1
                sim.later(done, mma_q)
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
            def publish(a=acc, t=t):
                assert acc_busy[a] == 0
                for c in range(2):
                    accum[c][a] = t
                    tfull[c][a].arrive()
            sim.later(publish, mma_q)
            acc += 1
            if acc == 2:
                acc, acc_phase = 0, acc_phase ^ 1

    def epilogue(c):
        acc, acc_phase = 0, 0
        for t in range(ntiles):
            yield lambda a=acc, p=acc_phase: tfull[c][a].passed(p)
            assert accum[c][acc] == t, "epilogue read a stale accumulator"
            seen[c].append(t)
            yield None                                      # reading tensor memory takes a while
            assert accum[c][acc] == t, "accumulator overwritten while the epilogue read it"
            tempty[acc].arrive(128)                         # this CTA's 128 epilogue threads -> the leader's barrier
            acc += 1
            if acc == 2:
                acc, acc_phase = 0, acc_phase ^ 1

    for c in range(2):
        sim.spawn(producer(c))
        sim.spawn(epilogue(c))
    sim.spawn(mma())
    sim.run([tma_q[0], tma_q[1], mma_q])
    assert seen[0] == seen[1] == list(range(ntiles))


@pytest.mark.parametrize("stages,nkb,ntiles", [(6, 12, 7), (4, 12, 5), (2, 3, 9), (6, 1, 20), (3, 16, 4)])
def test_pair_kernel_protocol(stages, nkb, ntiles):
    for seed in range(60):
        run_pair(seed, stages, nkb, ntiles)


def test_pair_model_catches_a_wrong_arrival_count():
    """The model is not vacuous: with the leader expecting only its OWN bytes the MMA reads half-filled slots."""
    global_failures = 0
    for seed in range(40):
        try:
            run_pair_broken(seed)
        except AssertionError:
            global_failures += 1
    assert global_failures > 0


def run_pair_broken(seed):
    sim = Sim(seed)
    stages, nkb, ntiles, A_B = 3, 4, 4, 32768
    full = [MBar(1) for _ in range(stages)]
    empty = [[MBar(1) for _ in range(stages)] for _ in range(2)]
    slot = [[None] * stages for _ in range(2)]
    tma_q, mma_q = [[], []], []

    def producer(c):
        stage, phase = 0, 0
        for t in range(ntiles):
            for kb in range(nkb):
                yield lambda s=stage, p=phase: empty[c][s].passed(p ^ 1)
                if c == 0:
                    full[stage].arrive_expect_tx(A_B)       # BUG under test: only the leader's bytes
                def land(s=stage, t=t, kb=kb):
                    slot[c][s] = (t, kb)
                    full[s].complete_tx(A_B if c == 0 else 0)
                sim.later(land, tma_q[c])
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    def mma():
        stage, phase = 0, 0
        for t in range(ntiles):
            for kb in range(nkb):
                yield lambda s=stage, p=phase: full[s].passed(p)
                assert slot[0][stage] == (t, kb) and slot[1][stage] == (t, kb)
                def done(s=stage):
                    for c in range(2):
                        empty[c][s].arrive()
                sim.later(done, mma_q)
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1

    sim.spawn(producer(0)); sim.spawn(producer(1)); sim.spawn(mma())
    sim.run([tma_q[0], tma_q[1], mma_q])


# ------------------------------------------------------------------------------------------------
# gemm_rows_seeded_kernel: G CTAs; producer / MMA stream [na sample tiles] + [all tiles]; the epilogue runs phase A,
# a grid-wide barrier, then phase B
# ------------------------------------------------------------------------------------------------
def run_seeded(seed, grid, stages, nkb, tiles_per_cta, sample_tiles):
    sim = Sim(seed)
    counter = [0]
    target = grid
    written = [False] * grid
    queues = []
    done_tiles = [[] for _ in range(grid)]

    def make_cta(g):
        ntiles = tiles_per_cta[g]
        na = min(sample_tiles, ntiles)
        seq = [("A", t) for t in range(na)] + [("B", t) for t in range(ntiles)]
        full = [MBar(1) for _ in range(stages)]
        empty = [MBar(1) for _ in range(stages)]
        tfull = [MBar(1) for _ in range(2)]
        tempty = [MBar(128) for _ in range(2)]
        slot = [None] * stages
        accum = [None, None]
        tma_q, mma_q = [], []
        queues.extend([tma_q, mma_q])

        def producer():
            stage, phase = 0, 0
            for item in seq:
                for kb in range(nkb):
                    yield lambda s=stage, p=phase: empty[s].passed(p ^ 1)
                    full[stage].arrive_expect_tx(1)
                    def land(s=stage, item=item, kb=kb):
                        slot[s] = (item, kb)
                        full[s].complete_tx(1)
                    sim.later(land, tma_q)
                    stage += 1
                    if stage == stages:
                        stage, phase = 0, phase ^ 1

        def mma():
            stage, phase, acc, acc_phase = 0, 0, 0, 0
            for item in seq:
                yield lambda a=acc, p=acc_phase: tempty[a].passed(p ^ 1)
                for kb in range(nkb):
                    yield lambda s=stage, p=phase: full[s].passed(p)
                    assert slot[stage] == (item, kb)
                    sim.later(lambda s=stage: empty[s].arrive(), mma_q)
                    stage += 1
                    if stage == stages:
                        stage, phase = 0, phase ^ 1
                def publish(a=acc, item=item):
                    accum[a] = item
                    tfull[a].arrive()
                sim.later(publish, mma_q)
                acc += 1
                if acc == 2:
                    acc, acc_phase = 0, acc_phase ^ 1

        def epilogue():
            acc, acc_phase = 0, 0
            for t in range(sample_tiles):                   # phase A
                if t < na:
                    yield lambda a=acc, p=acc_phase: tfull[a].passed(p)
                    assert accum[acc] == ("A", t)
                    tempty[acc].arrive(128)
                    acc += 1
                    if acc == 2:
                        acc, acc_phase = 0, acc_phase ^ 1
            written[g] = True
            counter[0] += 1                                 # release-add after the block maxima are written
            yield lambda: counter[0] >= target              # spin
            assert all(written), "thresholds computed before every CTA wrote its block maxima"
            for t in range(ntiles):                         # phase B
                yield lambda a=acc, p=acc_phase: tfull[a].passed(p)
                assert accum[acc] == ("B", t)
                done_tiles[g].append(t)
                tempty[acc].arrive(128)
                acc += 1
                if acc == 2:
                    acc, acc_phase = 0, acc_phase ^ 1

        sim.spawn(producer()); sim.spawn(mma()); sim.spawn(epilogue())

    for g in range(grid):
        make_cta(g)
    sim.run(queues)
    assert all(done_tiles[g] == list(range(tiles_per_cta[g])) for g in range(grid))


@pytest.mark.parametrize("grid,stages,nkb,tiles,sample", [(4, 4, 12, [5, 5, 5, 4], 1), (3, 4, 3, [9, 1, 2], 2), (5, 3, 2, [1, 1, 1, 1, 1], 4),
                                                           (2, 4, 12, [33, 32], 2), (6, 2, 1, [3, 0, 3, 3, 3, 3], 1)])
def test_seeded_kernel_protocol(grid, stages, nkb, tiles, sample):
    for seed in range(40):
        run_seeded(seed, grid, stages, nkb, tiles, sample)
