"""CPU: discrete-event model of the synchronisation protocol of K1r (emap_b200/csrc/mlp_rg.cu).

The kernel could not be run on hardware in the round it was written, so its barrier protocol is
transcribed here role by role (producer warp, MMA-issuing warp, 16 epilogue warps) with the SAME barrier
counts, wait parities and use-count formulas as the CUDA source, and executed under randomised latencies
with mbarrier phase-parity semantics.  The model raises on
  * deadlock (every role blocked, nothing in flight),
  * a wait that passes on an aliased phase (detected through the data hazards below),
  * ring stage overwritten while an MMA still reads it / consumed before its copy landed,
  * A-tile chunk written while an MMA reading it is in flight, or an MMA reading the wrong version,
  * TMEM accumulator overwritten before all 16 epilogue warps have read it, or read before complete,
  * the positional-encoding image of a tile (global scratch, double buffered, written by all 16 epilogue warps
    during the previous tile's reverse steps) copied into chunk 0 by the TMA engine before it is complete,
    overwritten while a copy reads it, or the copy landing in chunk 0 under a running MMA,
  * (training mode) the TMA store of a finished A-tile chunk to the backward's stash -- issued by the 19th warp on
    st_ready, released by st_done -- reading a chunk that an epilogue warp or a PE copy overwrites under it, or a
    chunk that does not hold the layer it is meant to store.
It does not model the arithmetic (tests/test_rg_emulation.py does) nor PTX memory-ordering fences.
"""
import heapq
import random

import pytest

K_STAGES = {3: 3, 1: 4}
K_STEPS = 16             # forward layers 0..7, reverse layers 7..0 (the output layer is not an MMA step)
SKIP = 4
USES = (8, 8)            # accumulator uses per tile: buf 0 / buf 1   (kUsesPerBuf)
A_PER_TILE = 15          # kAPerTile
EPI_WARPS = 16


def step_nkc(s):
    return 1 if s == 0 else (5 if s == SKIP else 4)


class Hazard(AssertionError):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed, self.tx = name, count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.completed += 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, f"too many arrivals on {self.name}"
        self.tx += tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, n):
        self.tx -= n
        assert self.tx >= 0
        self._maybe_complete()

    def test(self, parity):
        """mbarrier.try_wait.parity: true iff the phase with this parity has completed, i.e. the phase in
        progress has the other parity.  (A barrier two phases ahead of its waiter aliases -- on purpose.)"""
        return (self.completed & 1) != (parity & 1)


class Sim:
    def __init__(self, nterms, iters, seed, split_tail=False, training=True):
        self.rng = random.Random(seed)
        self.nterms, self.iters = nterms, iters
        self.training = training
        self.st_ready = [MBar(f"st_ready{i}", EPI_WARPS) for i in range(4)]
        self.st_done = [MBar(f"st_done{i}", 1) for i in range(4)]
        self.store_readers = [0] * 4                                # TMA stash stores in flight reading chunk c
        self.split_tail = split_tail and nterms == 3
        self.parts = 2 if nterms == 3 else 1
        ns = K_STAGES[nterms]
        self.ns = ns
        self.full = [MBar(f"full{i}", 1) for i in range(ns)]
        self.empty = [MBar(f"empty{i}", 1) for i in range(ns)]
        self.a_ready = [MBar(f"a_ready{i}", EPI_WARPS) for i in range(4)] + [MBar("a_ready4", 1)]
        self.pe_done = MBar("pe_done", EPI_WARPS)
        # tile schedule: completion k of sched_ready publishes the tile of iteration k in sched_tile[k & 1]
        # (True = a tile, False = past the end); completion 0 is made at barrier initialisation
        self.sched_ready = MBar("sched_ready", 1)
        self.sched_tile = [True, None]
        self.sched_ready.arrive()
        self.img_ver = [[None] * EPI_WARPS for _ in range(2)]     # PE image b: tile iteration each warp wrote
        self.img_readers = [0, 0]                                   # TMA copies in flight reading image b
        self.chunk0_copy = False                                    # a PE copy into chunk 0 is in flight
        self.acc_full = [[MBar(f"acc_full{b}{h}", 1) for h in range(2)] for b in range(2)]   # [buf][N half]
        self.acc_empty = [MBar("acc_empty0", EPI_WARPS), MBar("acc_empty1", EPI_WARPS)]
        self.c0_free = MBar("c0_free", 1)
        self.now = 0.0
        self.events = []          # (time, seq, fn)
        self.seq = 0
        # ---- resource state for the hazard checks
        self.stage_data = [None] * ns           # (iter, s, ic, part) landed in the stage
        self.stage_copying = [False] * ns
        self.stage_readers = [0] * ns           # MMAs in flight reading the stage
        # A chunk c: per-warp version tags + MMAs in flight reading it
        self.chunk_ver = [[None] * EPI_WARPS for _ in range(4)]
        self.chunk_readers = [0] * 4
        self.acc_ver = [[None, None], [None, None]]   # [buf][N half]: (iter, s) once complete
        self.acc_writing = [None, None]         # (iter, s) while MMAs in flight
        self.acc_reads_left = [0, 0]            # epilogue warps that still have to read the current version
        self.mma_queue = []                     # in-order completion of the tensor pipe
        self.mma_busy_until = 0.0
        self.blocked = {}
        self.done = 0

    # ------------------------------------------------------------------ event plumbing
    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    def lat(self, lo, hi):
        return self.rng.uniform(lo, hi)

    def run_role(self, name, gen):
        """advance a role until it blocks on a barrier or finishes"""
        try:
            while True:
                op = next(gen)
                if op[0] == "wait":
                    _, bar, parity = op
                    if not bar.test(parity):
                        self.blocked[name] = (gen, bar, parity)
                        return
                elif op[0] == "delay":
                    self.at(op[1], lambda n=name, g=gen: self.run_role(n, g))
                    return
        except StopIteration:
            self.done += 1

    def wake(self):
        for name in list(self.blocked):
            gen, bar, parity = self.blocked[name]
            if bar.test(parity):
                del self.blocked[name]
                # a woken role resumes after a small random latency (warp scheduling)
                self.at(self.lat(0.0, 0.3), lambda n=name, g=gen: self.run_role(n, g))

    def run(self):
        roles = {"producer": self.producer(), "issuer": self.issuer()}
        if self.training:
            roles["stash_io"] = self.stash_io()
        for w in range(EPI_WARPS):
            roles[f"epi{w}"] = self.epilogue(w)
        for name, gen in roles.items():
            self.at(self.lat(0, 1), lambda n=name, g=gen: self.run_role(n, g))
        nroles = len(roles)
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            self.wake()
        if self.done != nroles:
            state = {n: (b.name, p, b.completed) for n, (_, b, p) in self.blocked.items()}
            raise Hazard(f"deadlock: {state}")

    # ------------------------------------------------------------------ asynchronous agents
    def bulk_copy(self, stage, tag):
        if self.stage_readers[stage]:
            raise Hazard(f"copy into ring stage {stage} while an MMA reads it ({tag})")
        if self.stage_copying[stage]:
            raise Hazard(f"two copies in flight into ring stage {stage}")
        self.stage_copying[stage] = True
        self.stage_data[stage] = None

        def land():
            self.stage_copying[stage] = False
            self.stage_data[stage] = tag
            self.full[stage].complete_tx(1)
        self.at(self.lat(0.5, 6.0), land)

    def pe_copy(self, it, s, b):
        """TMA bulk copy of PE image b (tile iteration `it`) into A chunk 0, as the PE operand of step s"""
        for w in range(EPI_WARPS):
            if self.img_ver[b][w] != it:
                raise Hazard(f"PE copy for {(it, s)} reads image {b}: warp {w} wrote {self.img_ver[b][w]}")
        if self.chunk_readers[0]:
            raise Hazard(f"PE copy for {(it, s)} issued into chunk 0 under {self.chunk_readers[0]} MMAs in flight")
        if self.store_readers[0]:
            raise Hazard(f"PE copy for {(it, s)} issued into chunk 0 under a stash store reading it")
        if self.chunk0_copy:
            raise Hazard("two PE copies in flight")
        self.chunk0_copy = True
        self.img_readers[b] += 1

        def land():
            if self.chunk_readers[0]:
                raise Hazard(f"PE copy for {(it, s)} lands in chunk 0 under a running MMA")
            self.chunk0_copy = False
            self.img_readers[b] -= 1
            for w in range(EPI_WARPS):
                self.chunk_ver[0][w] = ("PE", it, s)
            self.a_ready[4].complete_tx(1)
        self.at(self.lat(0.5, 6.0), land)

    def mma_group(self, it, s, ic, parts, chunk, buf, first):
        """the MMAs of one K chunk over the given [(part, ring stage)]: read ring stages + A chunk,
        accumulate into TMEM buf"""
        for part, stage in parts:
            if self.stage_data[stage] != (it, s, ic, part):
                raise Hazard(f"MMA {(it, s, ic, part)} reads ring stage {stage} holding {self.stage_data[stage]}")
        want = ("PE", it, s) if chunk == 4 else (it, s)
        phys = 0 if chunk == 4 else chunk
        for w in range(EPI_WARPS):
            if self.chunk_ver[phys][w] != want:
                raise Hazard(f"MMA {(it, s, ic)} reads chunk {phys}: warp {w} wrote {self.chunk_ver[phys][w]}, want {want}")
        if first:
            if self.acc_reads_left[buf]:
                raise Hazard(f"step {(it, s)} overwrites TMEM buf {buf} with {self.acc_reads_left[buf]} reads outstanding")
            self.acc_writing[buf] = (it, s)
            self.acc_ver[buf] = [None, None]
        elif self.acc_writing[buf] != (it, s):
            raise Hazard(f"accumulating step {(it, s)} into buf {buf} owned by {self.acc_writing[buf]}")
        for _, stage in parts:
            self.stage_readers[stage] += 1
        self.chunk_readers[phys] += 1
        start = max(self.now, self.mma_busy_until)
        self.mma_busy_until = start + self.lat(0.5, 2.0)

        def fin():
            for _, stage in parts:
                self.stage_readers[stage] -= 1
            self.chunk_readers[phys] -= 1
        self.mma_queue.append((self.mma_busy_until, fin))
        self.at(self.mma_busy_until - self.now, self._retire)

    def _retire(self):
        while self.mma_queue and self.mma_queue[0][0] <= self.now + 1e-12:
            self.mma_queue.pop(0)[1]()

    def commit(self, fn):
        """tcgen05.commit: fires once every MMA issued so far has completed"""
        t = max(self.mma_busy_until, self.now)
        self.at(t - self.now + 1e-9, fn)

    # ------------------------------------------------------------------ roles (mirroring mlp_rg.cu)
    def _schedule(self, it):
        """what every role does at the top of iteration `it` (after the wait): read the published slot"""
        v = self.sched_tile[it & 1]
        if v is None:
            raise Hazard(f"iteration {it}: schedule slot read before it was published")
        return v

    def producer(self):
        stage, rnd = 0, 0
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for s in range(K_STEPS):
                for ip in range(step_nkc(s) * 2):
                    if self.nterms == 1 and (ip & 1):
                        continue
                    if rnd > 0:
                        yield ("wait", self.empty[stage], (rnd - 1) & 1)
                    self.full[stage].arrive(tx=1)                   # arrive.expect_tx
                    self.bulk_copy(stage, (it, s, ip // 2, ip & 1))
                    yield ("delay", self.lat(0.05, 0.3))
                    stage += 1
                    if stage == self.ns:
                        stage, rnd = 0, rnd + 1

    def issuer(self):
        stage, rnd = 0, 0
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for s in range(K_STEPS):
                buf = s & 1
                started = it * USES[buf] + (s >> 1)
                if started > 0:
                    yield ("wait", self.acc_empty[buf], (started - 1) & 1)
                nkc = step_nkc(s)
                split = self.split_tail and s not in (0, SKIP, K_STEPS - 1)

                def half_done(buf=buf, it=it, s=s, half=0):
                    self.acc_ver[buf][half] = (it, s)
                    if self.acc_ver[buf][0] == (it, s) and self.acc_ver[buf][1] == (it, s):
                        self.acc_writing[buf] = None
                        self.acc_reads_left[buf] = EPI_WARPS      # every epilogue warp still has to drain it
                    self.acc_full[buf][half].arrive()
                for ic in range(nkc):
                    c = 4 if s == 0 else (ic if ic < 4 else 4)
                    uses = (it * 2 + (1 if s == SKIP else 0)) if c == 4 else (it * A_PER_TILE + (s - 1))
                    yield ("wait", self.a_ready[c], uses & 1)
                    if split and ic == nkc - 1:
                        sidx = []
                        for part in range(self.parts):
                            yield ("wait", self.full[stage], rnd & 1)
                            sidx.append((part, stage))
                            stage += 1
                            if stage == self.ns:
                                stage, rnd = 0, rnd + 1
                        for half in range(2):
                            self.mma_group(it, s, ic, sidx, c, buf, first=False)
                            self.commit(lambda h=half, f=half_done: f(half=h))
                        for _, st in sidx:
                            self.commit(lambda st=st: self.empty[st].arrive())
                        yield ("delay", self.lat(0.05, 0.4))
                        continue
                    for part in range(self.parts):
                        yield ("wait", self.full[stage], rnd & 1)
                        self.mma_group(it, s, ic, [(part, stage)], c, buf, first=(ic == 0 and part == 0))
                        st = stage
                        self.commit(lambda st=st: self.empty[st].arrive())
                        yield ("delay", self.lat(0.05, 0.4))
                        stage += 1
                        if stage == self.ns:
                            stage, rnd = 0, rnd + 1
                    if s == SKIP and ic == 0:
                        self.commit(self.c0_free.arrive)
                if not split:
                    self.commit(lambda f=half_done: (f(half=0), f(half=1)))
                yield ("delay", self.lat(0.02, 0.1))

    def stash_io(self):
        """19th warp (training): chunk c of h_{l+1} (forward layers 0..6) from the A tile to the stash, one TMA store"""
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for l in range(7):
                for c in range(4):
                    yield ("wait", self.st_ready[c], (it * 7 + l) & 1)
                    for w in range(EPI_WARPS):
                        if self.chunk_ver[c][w] != (it, l + 1):
                            raise Hazard(f"stash store {(it, l, c)} reads chunk {c}: warp {w} wrote {self.chunk_ver[c][w]}")
                    self.store_readers[c] += 1
                    yield ("delay", self.lat(0.3, 3.0))           # cp.async.bulk.wait_group.read 0
                    self.store_readers[c] -= 1
                    self.st_done[c].arrive()

    def _write_chunk(self, w, phys, tag):
        if self.chunk_readers[phys]:
            raise Hazard(f"warp {w} writes chunk {phys} ({tag}) under {self.chunk_readers[phys]} MMAs in flight")
        if self.store_readers[phys]:
            raise Hazard(f"warp {w} writes chunk {phys} ({tag}) under a stash store reading it")
        if phys == 0 and self.chunk0_copy:
            raise Hazard(f"warp {w} writes chunk 0 ({tag}) while a PE copy into it is in flight")
        self.chunk_ver[phys][w] = tag

    def _encode(self, w, it):
        """this warp's part of the PE image of tile iteration `it` -> global scratch image it & 1"""
        b = it & 1
        if self.img_readers[b]:
            raise Hazard(f"warp {w} overwrites PE image {b} (tile {it}) under a copy in flight")
        self.img_ver[b][w] = it
        self.pe_done.arrive()

    def _read_acc(self, w, buf, it, s, half):
        if self.acc_ver[buf][half] != (it, s):
            raise Hazard(f"warp {w} reads TMEM buf {buf} half {half}: holds {self.acc_ver[buf][half]} "
                         f"(writing {self.acc_writing[buf]}), want {(it, s)}")

    def epilogue(self, w):
        sub = w >> 2
        # prologue: encode tile 0; warp 0 hands the image to the TMA engine
        yield ("delay", self.lat(0.2, 1.5))
        self._encode(w, 0)
        if w == 0:
            if not getattr(self, "skip_pe_wait", False):
                yield ("wait", self.pe_done, 0)
            self.a_ready[4].arrive(tx=1)
            self.pe_copy(0, 0, 0)
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            # forward layers 0..7
            for l in range(8):
                buf = l & 1
                par = (it * USES[buf] + (l >> 1)) & 1
                yield ("wait", self.acc_full[buf][0], par)
                if l == getattr(self, "publish_layer", 2) and w == 0:   # the scheduler thread publishes iteration it + 1
                    self.sched_tile[(it + 1) & 1] = (it + 1 < self.iters)
                    self.sched_ready.arrive()
                for chunk in range(4):
                    if chunk == 2:
                        yield ("wait", self.acc_full[buf][1], par)
                    self._read_acc(w, buf, it, l, chunk >> 1)
                    if self.training and l >= 1 and not getattr(self, "skip_st_done_wait", False):
                        yield ("wait", self.st_done[chunk], (it * 7 + l - 1) & 1)
                    yield ("delay", self.lat(0.1, 1.0))
                    self._write_chunk(w, chunk, (it, l + 1))
                    self.a_ready[chunk].arrive()
                    if self.training and l < 7:
                        self.st_ready[chunk].arrive()
                self.acc_reads_left[buf] -= 1
                self.acc_empty[buf].arrive()
                if l == SKIP - 1 and w == 0:
                    yield ("wait", self.c0_free, it & 1)
                    if self.training and not getattr(self, "skip_st_done_wait", False):
                        yield ("wait", self.st_done[0], (it * 7 + 3) & 1)
                    self.a_ready[4].arrive(tx=1)
                    self.pe_copy(it, SKIP, it & 1)
            # (layer 7's epilogue above wrote the sweep's seed alpha_7 as the input of step 8)
            # reverse steps 8..14
            for s in range(8, 15):
                buf = s & 1
                par = (it * USES[buf] + (s >> 1)) & 1
                yield ("wait", self.acc_full[buf][0], par)
                for chunk in range(4):
                    if chunk == 2:
                        yield ("wait", self.acc_full[buf][1], par)
                    self._read_acc(w, buf, it, s, chunk >> 1)
                    yield ("delay", self.lat(0.1, 1.0))
                    self._write_chunk(w, chunk, (it, s + 1))
                    self.a_ready[chunk].arrive()
                self.acc_reads_left[buf] -= 1
                self.acc_empty[buf].arrive()
                if s == 8:                                # the next tile's PE, in the MMA-bound reverse steps
                    yield ("wait", self.sched_ready, (it + 1) & 1)
                    if self._schedule(it + 1):
                        yield ("delay", self.lat(0.2, 1.5))
                        self._encode(w, it + 1)
            # step 15: PE adjoint -> gradient; the exchange slots live in chunk 3 of the (dead) A tile
            yield ("wait", self.acc_full[1][0], (it * USES[1] + 7) & 1)
            if w == 0 and self._schedule(it + 1):         # chunk 0 is free: next tile's PE image on its way
                if not getattr(self, "skip_pe_wait", False):
                    yield ("wait", self.pe_done, (it + 1) & 1)
                self.a_ready[4].arrive(tx=1)
                self.pe_copy(it + 1, 0, (it + 1) & 1)
            self._read_acc(w, 1, it, 15, 0)
            self.acc_reads_left[1] -= 1
            self.acc_empty[1].arrive()
            yield ("delay", self.lat(0.1, 0.8))
            self._write_chunk(w, 3, ("slots", it))
            yield ("delay", self.lat(0.05, 0.5))     # (named barriers of the lane quarter: no mbarrier involved)


@pytest.mark.parametrize("nterms,split", [(3, False), (3, True), (1, False)])
def test_protocol_no_deadlock_no_hazard(nterms, split):
    for seed in range(40):
        Sim(nterms, iters=3, seed=seed, split_tail=split).run()
        Sim(nterms, iters=3, seed=seed, split_tail=split, training=False).run()


def _heavy_tailed(self, lo, hi):
    """mostly in range, sometimes 30x slower or 30x faster: very different regimes of who waits for whom"""
    r, base = self.rng.random(), self.rng.uniform(lo, hi)
    return base * 30 if r < 0.1 else (base / 30 + 1e-6 if r < 0.2 else base)


@pytest.mark.parametrize("nterms,split", [(3, False), (3, True), (1, False)])
def test_protocol_under_heavy_tailed_latencies(nterms, split, monkeypatch):
    monkeypatch.setattr(Sim, "lat", _heavy_tailed)
    for seed in range(300, 340):
        Sim(nterms, iters=4, seed=seed, split_tail=split).run()


def test_model_detects_a_pe_copy_before_the_image_is_complete():
    """the PE hand-off has teeth too: without the pe_done wait the first TMA copy reads a half-written image
    (later tiles are also covered by the data dependencies of the sweep -- every warp's step-9 epilogue comes
    after its encoding -- there pe_done supplies the release/acquire ordering of the image's global stores)"""
    caught = 0
    for seed in range(20):
        sim = Sim(3, iters=2, seed=seed)
        sim.skip_pe_wait = True
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught > 0


def test_model_detects_a_chunk_overwritten_under_its_stash_store(monkeypatch):
    """without the st_done waits a slow TMA store still reads the chunk that the next layer's epilogue (or the skip
    layer's PE copy) overwrites"""
    monkeypatch.setattr(Sim, "lat", _heavy_tailed)
    caught = 0
    for seed in range(40):
        sim = Sim(3, iters=3, seed=seed)
        sim.skip_st_done_wait = True
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught > 0


def test_model_detects_a_schedule_published_too_early(monkeypatch):
    """publishing the next tile before every role has consumed the current completion (during step 0 instead of
    step 2) lets a straggler wait on an aliased parity: caught as a deadlock"""
    monkeypatch.setattr(Sim, "lat", _heavy_tailed)
    caught = 0
    for seed in range(60):
        sim = Sim(3, iters=4, seed=seed)
        sim.publish_layer = 0
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught > 0


def test_model_detects_a_wrong_parity():
    """the model has teeth: an off-by-one in the a_ready use count must be caught"""
    global A_PER_TILE
    old = A_PER_TILE
    A_PER_TILE = 16
    try:
        with pytest.raises(AssertionError):
            for seed in range(10):
                Sim(3, iters=3, seed=seed).run()
    finally:
        A_PER_TILE = old


def test_model_detects_a_wrong_accumulator_count():
    """each buffer is used 8 times per tile, not 9: the wrong count must be caught from the second tile on"""
    global USES
    old = USES
    USES = (9, 8)
    try:
        with pytest.raises(AssertionError):
            for seed in range(10):
                Sim(3, iters=3, seed=seed).run()
    finally:
        USES = old


def test_model_constants_match_the_cuda_source():
    """the model is a transcription: at least its constants are read back from mlp_rg.cu"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "emap_b200", "csrc", "mlp_rg.cu")).read()

    def const(name):
        m = re.search(rf"constexpr\s+\w+\s+{name}\s*=\s*(\d+)", src)
        assert m, name
        return int(m.group(1))
    assert const("kSteps") == K_STEPS
    assert const("kUsesPerBuf") == USES[0] == USES[1]
    assert const("kAPerTile") == A_PER_TILE
    assert re.search(r"kStages = \(NTERMS == 3\) \? 3 : 4;", src) and K_STAGES == {3: 3, 1: 4}
    assert re.search(r"mbar_init\(&a_ready\[c\], kEpiWarps\)", src) and re.search(r"mbar_init\(&a_ready\[4\], 1\)", src)
    assert re.search(r"mbar_init\(pe_done, kEpiWarps\)", src)
    assert "mbar_wait(pe_done, (uint32_t)(iter + 1) & 1" in src and "mbar_wait(pe_done, 0" in src
    assert re.search(r"mbar_init\(sched_ready, 1\)", src) and src.count("mbar_wait(sched_ready, (uint32_t)iter & 1") == 4
    assert "if (l == 2 && scheduler)" in src and "mbar_wait(sched_ready, (uint32_t)(iter + 1) & 1" in src
    assert re.search(r"mbar_init\(&acc_empty\[b\], kEpiWarps\)", src)
    assert "mbar_init(&st_ready[c], kEpiWarps); mbar_init(&st_done[c], 1)" in src
    assert "mbar_wait(&st_ready[c], ((uint32_t)iter * 7u + (uint32_t)l) & 1" in src
    assert "mbar_wait(&st_done[chunk], ((uint32_t)iter * 7u + (uint32_t)(l - 1)) & 1" in src
    assert "mbar_wait(&st_done[0], ((uint32_t)iter * 7u + 3u) & 1" in src
    assert "uses & 1" in src and "(uint32_t)iter * kAPerTile + (uint32_t)(s - 1)" in src
    assert "(uint32_t)iter * 2u + (s == kSkipLayer ? 1u : 0u)" in src
    assert "(uint32_t)iter * kUsesPerBuf + (uint32_t)(s >> 1)" in src
    dev = open(os.path.join(root, "emap_b200", "csrc", "mlp_dev.cuh")).read()
    assert re.search(r"constexpr int kEpiWarps = (\d+);", dev).group(1) == str(EPI_WARPS)
