"""CPU: discrete-event model of the tangent forward with TMA-staged stash rows (mlp_kernel<1, MODE 3, __half, 1> in
emap_b200/csrc/mlp_tc.cu): producer, MMA issuer, 16 epilogue warps, the stash I/O warp.  Built on the plumbing of
tests/test_rev_protocol.py; what is specific here: the positional-encoding chunk that lives in activation chunk 0
(written at tile start and again for the skip term of layer 4 behind `c0_free`), one K chunk for layer 0 and five for
layer 4, the value rows h_{l+1} arriving in the slots a layer ahead (`u_full`) and the tangent rows leaving from them
(`out_done`), eight of each per tile.  Raises on deadlock and on the data hazards listed in test_rev_protocol.py plus:
the PE written into chunk 0 under an MMA that still reads it, an MMA reading chunk 0 with the wrong content."""
import pytest

from tests.test_rev_protocol import RevSim
from tests.test_rg_protocol import Hazard, MBar

EPI_WARPS = 16
PE_WARPS = 8             # warps with sub < 2 write the PE chunk (mbar_init(&a_ready[4], 8 * kArr))
K_STAGES = 3             # SmemPlan<1, 3, false, true>::kStages
SKIP = 4
A_PER_TILE = 7           # ModeInfo<3>::kAPerTile
USES = (4, 4)            # accumulator uses per tile: buf 0 (kUses0) / buf 1


def nkc(l):
    return 1 if l == 0 else (5 if l == SKIP else 4)


class TanSim(RevSim):
    def __init__(self, iters, seed, heavy=False):
        super().__init__(iters, seed, heavy)
        self.a_ready.append(MBar("a_ready4", PE_WARPS))
        self.c0_free = MBar("c0_free", 1)
        self.pe_ver = [None] * EPI_WARPS                 # chunk 0 as PE operand: (it, layer) per PE warp
        self.chunk0_is_pe = False

    def run(self):
        self.roles = {"producer": self.producer(), "issuer": self.issuer(), "io": self.io()}
        for w in range(EPI_WARPS):
            self.roles[f"epi{w}"] = self.epilogue(w)
        import heapq
        for name, gen in self.roles.items():
            self.at(self.lat(0, 1), lambda n=name, g=gen: self.run_role(n, g))
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            self.wake()
        if self.done != len(self.roles):
            raise Hazard(f"deadlock: { {n: (b.name, p, b.completed) for n, (_, b, p) in self.blocked.items()} }")

    # ---- MMA of one K chunk: c < 4 = activation chunk written by epilogue(l-1); c == 4 = PE in chunk 0
    def mma(self, it, l, ic, c, stage, buf):
        if self.stage_data[stage] != (it, l, ic):
            raise Hazard(f"MMA {(it, l, ic)} reads ring stage {stage} holding {self.stage_data[stage]}")
        phys = 0 if c == 4 else c
        if c == 4:
            for w in range(PE_WARPS):
                if self.pe_ver[w] != (it, l) or not self.chunk0_is_pe:
                    raise Hazard(f"MMA {(it, l, ic)} wants the PE in chunk 0: warp {w} wrote {self.pe_ver[w]}, "
                                 f"chunk 0 holds {'PE' if self.chunk0_is_pe else 'activations'}")
        else:
            if c == 0 and self.chunk0_is_pe:
                raise Hazard(f"MMA {(it, l, ic)} reads activations from chunk 0, which holds the PE")
            for w in range(EPI_WARPS):
                if self.chunk_ver[c][w] != (it, l - 1):
                    raise Hazard(f"MMA {(it, l, ic)} reads chunk {c}: warp {w} wrote {self.chunk_ver[c][w]}")
        if ic == 0:
            if self.acc_reads_left[buf]:
                raise Hazard(f"layer {(it, l)} overwrites TMEM buf {buf} with reads outstanding")
            self.acc_writing[buf], self.acc_ver[buf] = (it, l), None
        self.stage_readers[stage] += 1
        self.chunk_readers[phys] += 1
        start = max(self.now, self.mma_busy_until)
        self.mma_busy_until = start + self.lat(0.3, 1.5)

        def fin():
            self.stage_readers[stage] -= 1
            self.chunk_readers[phys] -= 1
        self.mma_queue.append((self.mma_busy_until, fin))
        self.at(self.mma_busy_until - self.now, self._retire)

    def producer(self):
        stage, rnd, it = 0, 0, -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            for l in range(8):
                for ic in range(nkc(l)):                  # single-MMA mode: the hi image of each K chunk only
                    if rnd > 0:
                        yield ("wait", self.empty[stage], (rnd - 1) & 1)
                    self.full[stage].arrive(tx=1)
                    self.ring_copy(stage, (it, l, ic))
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
            for l in range(8):
                buf = l & 1
                started = it * USES[buf] + (l >> 1)
                if started > 0:
                    yield ("wait", self.acc_empty[buf], (started - 1) & 1)
                for ic in range(nkc(l)):
                    c = 4 if l == 0 else (ic if ic < 4 else 4)
                    uses = (it * 2 + (1 if l == SKIP else 0)) if c == 4 else (it * A_PER_TILE + (l - 1))
                    yield ("wait", self.a_ready[c], uses & 1)
                    yield ("wait", self.full[stage], rnd & 1)
                    self.mma(it, l, ic, c, stage, buf)
                    st = stage
                    self.commit(lambda st=st: self.empty[st].arrive())
                    yield ("delay", self.lat(0.05, 0.4))
                    stage += 1
                    if stage == K_STAGES:
                        stage, rnd = 0, rnd + 1
                    if l == SKIP and ic == 0:
                        self.commit(self.c0_free.arrive)

                def acc_done(buf=buf, it=it, l=l):
                    self.acc_ver[buf], self.acc_writing[buf] = (it, l), None
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
            for l in range(8):
                def refill(c, l=l, it=it):
                    if l < 7:
                        self.load_u(c, (it, l + 1))
                    elif nxt:
                        self.load_u(c, (it + 1, 0))
                if l == 7:
                    yield ("wait", self.sched_ready, (it + 1) & 1)
                    nxt = self._schedule(it + 1)
                pending = []
                for c in range(4):
                    yield ("wait", self.slot_done[c], (it * 8 + l) & 1)       # out_done
                    for w in range(EPI_WARPS):
                        if self.slot_out[c][w] != (it, l):
                            raise Hazard(f"TMA store {(it, l, c)} reads slot {c}: warp {w} wrote {self.slot_out[c][w]}")
                    self.slot_store_reading[c] = True
                    pending.append(c)
                    if c > 0:
                        yield ("delay", self.lat(0.1, 2.0))
                        done = pending.pop(0)
                        self.slot_store_reading[done] = False
                        refill(done)
                yield ("delay", self.lat(0.1, 2.0))
                self.slot_store_reading[pending.pop(0)] = False
                refill(3)

    def _write_pe(self, w, it, l):
        if self.chunk_readers[0]:
            raise Hazard(f"warp {w} writes the PE of {(it, l)} into chunk 0 under {self.chunk_readers[0]} MMAs in flight")
        self.pe_ver[w] = (it, l)
        self.chunk0_is_pe = True
        self.a_ready[4].arrive()

    def epilogue(self, w):
        sub = w >> 2
        it = -1
        while True:
            it += 1
            yield ("wait", self.sched_ready, it & 1)
            if not self._schedule(it):
                return
            if sub < 2:                                   # input stage: positional encoding (tangent) -> chunk 0
                yield ("delay", self.lat(0.2, 1.5))
                self._write_pe(w, it, 0)
            for l in range(8):
                buf = l & 1
                yield ("wait", self.acc_full[buf], (it * USES[buf] + (l >> 1)) & 1)
                if l == getattr(self, "publish_layer", 2) and w == 0:
                    self.sched_tile[(it + 1) & 1] = (it + 1 < self.iters)
                    self.sched_ready.arrive()
                for c in range(4):
                    if self.acc_ver[buf] != (it, l):
                        raise Hazard(f"warp {w} reads TMEM buf {buf}: holds {self.acc_ver[buf]}, want {(it, l)}")
                    yield ("wait", self.u_full[c], (it * 8 + l) & 1)
                    if self.slot_u[c] != (it, l):
                        raise Hazard(f"warp {w} reads slot {c}: holds h of {self.slot_u[c]}, want {(it, l)}")
                    self.slot_users[c] += 1
                    yield ("delay", self.lat(0.1, 1.0))
                    if l < 7:
                        if self.chunk_readers[c]:
                            raise Hazard(f"warp {w} writes A chunk {c} under {self.chunk_readers[c]} MMAs in flight")
                        self.chunk_ver[c][w] = (it, l)
                        if c == 0:
                            self.chunk0_is_pe = False
                    self.slot_out[c][w] = (it, l)
                    self.slot_users[c] -= 1
                    if l < 7:
                        self.a_ready[c].arrive()
                    self.slot_done[c].arrive()
                self.acc_reads_left[buf] -= 1
                self.acc_empty[buf].arrive()
                if l == SKIP - 1 and sub < 2:             # skip connection: the PE again, as the 5th K chunk of layer 4
                    if not getattr(self, "skip_c0_wait", False):
                        yield ("wait", self.c0_free, it & 1)
                    yield ("delay", self.lat(0.2, 1.5))
                    self._write_pe(w, it, SKIP)


def test_protocol_no_deadlock_no_hazard():
    for seed in range(40):
        TanSim(iters=3, seed=seed).run()


def test_protocol_under_heavy_tailed_latencies():
    for seed in range(300, 340):
        TanSim(iters=4, seed=seed, heavy=True).run()


def test_model_detects_the_pe_written_under_a_running_mma():
    """without the c0_free wait the skip layer's PE lands in chunk 0 while layer 4's first MMAs still read h_4 there"""
    caught = 0
    for seed in range(40):
        sim = TanSim(iters=2, seed=seed, heavy=True)
        sim.skip_c0_wait = True
        try:
            sim.run()
        except AssertionError:
            caught += 1
    assert caught > 0


def test_model_constants_match_the_cuda_source():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "emap_b200", "csrc", "mlp_tc.cu")).read()
    dev = open(os.path.join(root, "emap_b200", "csrc", "mlp_dev.cuh")).read()
    assert "static constexpr int kStages = ((NTERMS == 3 || kSlots) ? 3 : 4) * (PAIR ? 2 : 1);" in dev
    assert "static constexpr int kAPerTile = (MODE >= 2) ? 7 : 8;" in dev
    assert "static constexpr int kUses0 = (MODE >= 2) ? 4 : 5;" in dev
    assert "mbar_init(&a_ready[4], 8 * kArr);" in src
    assert "mbar_init(&u_full[c], 1); mbar_init(&out_done[c], kEpiWarps);" in src
    assert "mbar_wait(&out_done[c], ((uint32_t)iter * 8u + (uint32_t)l) & 1" in src
    assert "mbar_wait(&u_full[chunk], ((uint32_t)iter * 8u + (uint32_t)l) & 1" in src
    assert "(uint32_t)iter * 2u + (l == kSkipLayer ? 1u : 0u)" in src
    assert "(uint32_t)iter * (uint32_t)MI::kAPerTile + (uint32_t)(l - 1)" in src
    assert "mbar_wait(c0_free, (uint32_t)iter & 1, 520);" in src
    assert re.search(r"if \(l == 2 && scheduler\)", src)
