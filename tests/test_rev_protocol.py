"""CPU: discrete-event model of the synchronisation protocol of the reverse sweep (emap_b200/csrc/mlp_rev.cu).

Same method as tests/test_rg_protocol.py (whose mbarrier model it reuses): the kernel's five roles -- producer warp
(W^T ring), MMA-issuing warp, 16 epilogue warps, the I/O warp that owns the TMA traffic of both stashes -- are
transcribed with the SAME barrier counts, wait parities and use-count formulas as the CUDA source and executed under
randomised latencies with mbarrier phase-parity semantics.  The model raises on
  * deadlock,
  * a ring stage overwritten while an MMA reads it / consumed before its copy landed,
  * an A-tile chunk written while an MMA reading it is in flight, or an MMA reading the wrong version,
  * a TMEM accumulator overwritten before all 16 epilogue warps have read it, or read before complete,
  * a slot read by an epilogue warp before the TMA load of that (tile, stage) has landed, a TMA load landing in a
    slot that epilogue warps still read / rewrite or that a TMA store still reads, a TMA store reading a slot that
    does not hold the finished stage of all 16 warps.
It does not model the arithmetic nor PTX memory-ordering fences.
"""
import heapq
import random

import pytest

from tests.test_rg_protocol import Hazard, MBar

EPI_WARPS = 16
K_STAGES = 3             # kStagesT
REV_LAYERS = 7           # kRevLayers: MMA layers l = 7..1
STAGES = 8               # epilogue stages per tile (A_7 .. A_0)


class RevSim:
    def __init__(self, iters, seed, heavy=False):
        self.rng = random.Random(seed)
        self.heavy = heavy
        self.iters = iters
        self.full = [MBar(f"full{i}", 1) for i in range(K_STAGES)]
        self.empty = [MBar(f"empty{i}", 1) for i in range(K_STAGES)]
        self.a_ready = [MBar(f"a_ready{i}", EPI_WARPS) for i in range(4)]
        self.acc_full = [MBar("acc_full0", 1), MBar("acc_full1", 1)]
        self.acc_empty = [MBar("acc_empty0", EPI_WARPS), MBar("acc_empty1", EPI_WARPS)]
        self.u_full = [MBar(f"u_full{i}", 1) for i in range(4)]
        self.slot_done = [MBar(f"slot_done{i}", EPI_WARPS) for i in range(4)]
        self.sched_ready = MBar("sched_ready", 1)
        self.sched_tile = [True, None]
        self.sched_ready.arrive()
        self.now, self.events, self.seq = 0.0, [], 0
        self.blocked, self.done = {}, 0
        # resources
        self.stage_data = [None] * K_STAGES
        self.stage_copying = [False] * K_STAGES
        self.stage_readers = [0] * K_STAGES
        self.chunk_ver = [[None] * EPI_WARPS for _ in range(4)]     # A tile chunk c, per warp: (it, stage)
        self.chunk_readers = [0] * 4
        self.acc_ver = [None, None]
        self.acc_writing = [None, None]
        self.acc_reads_left = [0, 0]
        self.slot_u = [None] * 4                                    # (it, stage) whose U the slot holds (landed)
        self.slot_loading = [False] * 4
        self.slot_users = [0] * 4                                   # epilogue warps between their read and rewrite
        self.slot_out = [[None] * EPI_WARPS for _ in range(4)]      # per warp: (it, stage) whose A it wrote
        self.slot_store_reading = [False] * 4
        self.mma_queue, self.mma_busy_until = [], 0.0

    # ---- plumbing (as in test_rg_protocol.Sim)
    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    def lat(self, lo, hi):
        base = self.rng.uniform(lo, hi)
        if not self.heavy:
            return base
        r = self.rng.random()
        return base * 30 if r < 0.1 else (base / 30 + 1e-6 if r < 0.2 else base)

    def run_role(self, name, gen):
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
                self.at(self.lat(0.0, 0.3), lambda n=name, g=gen: self.run_role(n, g))

    def run(self):
        roles = {"producer": self.producer(), "issuer": self.issuer(), "io": self.io()}
        for w in range(EPI_WARPS):
            roles[f"epi{w}"] = self.epilogue(w)
        for name, gen in roles.items():
            self.at(self.lat(0, 1), lambda n=name, g=gen: self.run_role(n, g))
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            self.wake()
        if self.done != len(roles):
            raise Hazard(f"deadlock: { {n: (b.name, p, b.completed) for n, (_, b, p) in self.blocked.items()} }")

    def _schedule(self, it):
        v = self.sched_tile[it & 1]
        if v is None:
            raise Hazard(f"iteration {it}: schedule slot read before it was published")
        return v

    # ---- asynchronous agents
    def ring_copy(self, stage, tag):
        if self.stage_readers[stage] or self.stage_copying[stage]:
            raise Hazard(f"copy into ring stage {stage} while it is read / being written ({tag})")
        self.stage_copying[stage], self.stage_data[stage] = True, None

        def land():
            self.stage_copying[stage], self.stage_data[stage] = False, tag
            self.full[stage].complete_tx(1)
        self.at(self.lat(0.5, 6.0), land)

    def load_u(self, c, tag):
        """TMA loads of U for (tile iteration, stage) `tag` into slot c (arrive.expect_tx first)"""
        if self.slot_users[c]:
            raise Hazard(f"TMA load {tag} issued into slot {c} with {self.slot_users[c]} epilogue warps inside")
        if self.slot_store_reading[c]:
            raise Hazard(f"TMA load {tag} issued into slot {c} while a TMA store still reads it")
        if self.slot_loading[c]:
            raise Hazard(f"two TMA loads in flight into slot {c}")
        self.u_full[c].arrive(tx=1)
        self.slot_loading[c], self.slot_u[c] = True, None

        def land():
            if self.slot_users[c]:
                raise Hazard(f"TMA load {tag} lands in slot {c} under an epilogue warp")
            self.slot_loading[c], self.slot_u[c] = False, tag
            self.u_full[c].complete_tx(1)
        self.at(self.lat(0.5, 8.0), land)

    def mma_group(self, it, j, kc, stage, buf):
        if self.stage_data[stage] != (it, j, kc):
            raise Hazard(f"MMA {(it, j, kc)} reads ring stage {stage} holding {self.stage_data[stage]}")
        for w in range(EPI_WARPS):
            if self.chunk_ver[kc][w] != (it, j):            # A_{7-j}: written by epilogue stage j
                raise Hazard(f"MMA {(it, j, kc)} reads chunk {kc}: warp {w} wrote {self.chunk_ver[kc][w]}")
        if kc == 0:
            if self.acc_reads_left[buf]:
                raise Hazard(f"layer {(it, j)} overwrites TMEM buf {buf} with reads outstanding")
            self.acc_writing[buf], self.acc_ver[buf] = (it, j), None
        self.stage_readers[stage] += 1
        self.chunk_readers[kc] += 1
        start = max(self.now, self.mma_busy_until)
        self.mma_busy_until = start + self.lat(0.3, 1.5)

        def fin():
            self.stage_readers[stage] -= 1
            self.chunk_readers[kc] -= 1
        self.mma_queue.append((self.mma_busy_until, fin))
        self.at(self.mma_busy_until - self.now, self._retire)

    def _retire(self):
        while self.mma_queue and self.mma_queue[0][0] <= self.now + 1e-12:
            self.mma_queue.pop(0)[1]()

    def commit(self, fn):
        self.at(max(self.mma_busy_until, self.now) - self.now + 1e-9, fn)

    # ---- roles (mirroring mlp_rev.cu)
    def producer(self):
        stage, rnd, it = 0, 0, -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for i in range(REV_LAYERS * 4):
                if rnd > 0:
                    yield ("wait", self.empty[stage], (rnd - 1) & 1)
                self.full[stage].arrive(tx=1)
                self.ring_copy(stage, (it, i // 4, i % 4))
                yield ("delay", self.lat(0.05, 0.3))
                stage += 1
                if stage == K_STAGES:
                    stage, rnd = 0, rnd + 1

    def issuer(self):
        stage, rnd, it = 0, 0, -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for j in range(REV_LAYERS):
                buf = j & 1
                started = it * (3 if buf else 4) + (j >> 1)
                if started > 0:
                    yield ("wait", self.acc_empty[buf], (started - 1) & 1)
                for kc in range(4):
                    yield ("wait", self.a_ready[kc], (it * 7 + j) & 1)
                    yield ("wait", self.full[stage], rnd & 1)
                    self.mma_group(it, j, kc, stage, buf)
                    st = stage
                    self.commit(lambda st=st: self.empty[st].arrive())
                    yield ("delay", self.lat(0.05, 0.4))
                    stage += 1
                    if stage == K_STAGES:
                        stage, rnd = 0, rnd + 1

                def acc_done(buf=buf, it=it, j=j):
                    self.acc_ver[buf], self.acc_writing[buf] = (it, j), None
                    self.acc_reads_left[buf] = EPI_WARPS
                    self.acc_full[buf].arrive()
                self.commit(acc_done)

    def io(self):
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            if it == 0:
                for c in range(4):
                    self.load_u(c, (0, 0))
            nxt = False
            for s in range(STAGES):
                def refill(c, s=s, it=it):
                    if s < STAGES - 1:
                        self.load_u(c, (it, s + 1))
                    elif nxt:
                        self.load_u(c, (it + 1, 0))
                if s == STAGES - 1:
                    yield ("wait", self.sched_ready, (it + 1) & 1)
                    nxt = self._schedule(it + 1)
                pending = []
                for c in range(4):
                    yield ("wait", self.slot_done[c], (it * 8 + s) & 1)
                    for w in range(EPI_WARPS):
                        if self.slot_out[c][w] != (it, s):
                            raise Hazard(f"TMA store {(it, s, c)} reads slot {c}: warp {w} wrote {self.slot_out[c][w]}")
                    self.slot_store_reading[c] = True
                    pending.append(c)
                    if getattr(self, "skip_read_wait", False):
                        refill(c)                             # (teeth test: refill without wait_group.read)
                        pending.pop()
                        self.at(self.lat(0.1, 2.0), lambda c=c: self.slot_store_reading.__setitem__(c, False))
                    elif c > 0:                               # cp.async.bulk.wait_group.read 1
                        yield ("delay", self.lat(0.1, 2.0))
                        done = pending.pop(0)
                        self.slot_store_reading[done] = False
                        refill(done)
                if pending:
                    yield ("delay", self.lat(0.1, 2.0))       # wait_group.read 0
                    self.slot_store_reading[pending.pop(0)] = False
                    refill(3)

    def epilogue(self, w):
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for j in range(-1, REV_LAYERS):
                s, buf = j + 1, j & 1
                if j >= 0:
                    yield ("wait", self.acc_full[buf], (it * (3 if buf else 4) + (j >> 1)) & 1)
                    if j == getattr(self, "publish_stage", 1) and w == 0:
                        self.sched_tile[(it + 1) & 1] = (it + 1 < self.iters)
                        self.sched_ready.arrive()
                for c in range(4):
                    if j >= 0 and self.acc_ver[buf] != (it, j):
                        raise Hazard(f"warp {w} reads TMEM buf {buf}: holds {self.acc_ver[buf]}, want {(it, j)}")
                    if not getattr(self, "skip_u_wait", False):
                        yield ("wait", self.u_full[c], (it * 8 + s) & 1)
                    if self.slot_u[c] != (it, s):
                        raise Hazard(f"warp {w} reads slot {c}: holds U of {self.slot_u[c]}, want {(it, s)}")
                    self.slot_users[c] += 1
                    yield ("delay", self.lat(0.1, 1.0))
                    if s < STAGES - 1:                        # A_l, l >= 1: operand of the next MMA layer
                        if self.chunk_readers[c]:
                            raise Hazard(f"warp {w} writes A chunk {c} under {self.chunk_readers[c]} MMAs in flight")
                        self.chunk_ver[c][w] = (it, s)
                    self.slot_out[c][w] = (it, s)
                    self.slot_users[c] -= 1
                    if s < STAGES - 1:
                        self.a_ready[c].arrive()
                    self.slot_done[c].arrive()
                if j >= 0:
                    self.acc_reads_left[buf] -= 1
                    self.acc_empty[buf].arrive()


def test_protocol_no_deadlock_no_hazard():
    for seed in range(40):
        RevSim(iters=3, seed=seed).run()


def test_protocol_under_heavy_tailed_latencies():
    for seed in range(300, 340):
        RevSim(iters=4, seed=seed, heavy=True).run()


def test_model_detects_a_slot_read_before_its_load():
    """the slot protocol has teeth: without the u_full wait an epilogue warp reads a slot whose load is in flight"""
    caught = 0
    for seed in range(30):
        sim = RevSim(iters=2, seed=seed, heavy=True)
        sim.skip_u_wait = True
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught > 0


def test_model_detects_a_refill_before_the_store_has_read_the_slot():
    caught = 0
    for seed in range(10):
        sim = RevSim(iters=2, seed=seed)
        sim.skip_read_wait = True
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught == 10


def test_model_detects_a_schedule_that_is_never_published():
    caught = 0
    for seed in range(60):
        sim = RevSim(iters=4, seed=seed, heavy=True)
        sim.publish_stage = -5            # never published -> every role blocks at the top of the next iteration
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught == 60


def test_model_constants_match_the_cuda_source():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "emap_b200", "csrc", "mlp_rev.cu")).read()
    assert re.search(r"constexpr int kStagesT = (\d+);", src).group(1) == str(K_STAGES)
    assert re.search(r"constexpr int kRevLayers = (\d+);", src).group(1) == str(REV_LAYERS)
    assert "mbar_init(&u_full[c], 1);" in src and "mbar_init(&slot_done[c], kEpiWarps);" in src
    assert "mbar_init(&a_ready[c], kEpiWarps);" in src
    assert "mbar_wait(&slot_done[c], ((uint32_t)iter * 8u + (uint32_t)s) & 1" in src
    assert "const uint32_t io_par = ((uint32_t)iter * 8u + (uint32_t)(j + 1)) & 1;" in src
    assert "mbar_wait(&u_full[chunk], io_par" in src
    assert "mbar_wait(&a_ready[kc], ((uint32_t)iter * 7u + (uint32_t)j) & 1" in src
    assert "(uint32_t)iter * (buf ? 3u : 4u) + (uint32_t)(j >> 1)" in src
    assert "if (j == 1 && scheduler)" in src and src.count("mbar_wait(sched_ready, (uint32_t)iter & 1") == 4
    assert "mbar_wait(sched_ready, (uint32_t)(iter + 1) & 1" in src
    assert "bulk_wait_read1();" in src and "bulk_wait_read0();" in src
