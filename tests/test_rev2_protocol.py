"""CPU: discrete-event model of the two-tiles-in-flight reverse sweep (emap_b200/csrc/mlp_rev2.cu).

Same method as tests/test_rg_protocol.py (whose event plumbing is reused): the kernel's barrier protocol
-- counts, wait parities, use-count formulas, the X(j), Y(j), X(j+1), ... order of the MMA-issuing warp and
of the 16 epilogue warps, the 3-stage ring streamed per (tile, layer) -- transcribed role by role and run
under randomised latencies.  Raises on deadlock, ring / A-tile / TMEM hazards and phase aliasing.
"""
import pytest

from tests.test_rg_protocol import EPI_WARPS, Hazard, MBar, Sim

K_STAGES = 3
REV_LAYERS = 7          # MMA steps j = 0..6 (layers 7..1); epilogue stages j = -1..6
TILES = 2
A_PER_TILE = 7          # a_ready completions per tile slot and iteration (stages j = -1..5)


class Rev2Sim(Sim):
    def __init__(self, iters, seed, a_per_tile=A_PER_TILE):
        super().__init__(1, iters, seed)
        self.a_per_tile = a_per_tile
        self.ns = K_STAGES
        self.full = [MBar(f"full{i}", 1) for i in range(K_STAGES)]
        self.empty = [MBar(f"empty{i}", 1) for i in range(K_STAGES)]
        self.a_ready = [MBar(f"a_ready{t}{c}", EPI_WARPS) for t in range(TILES) for c in range(4)]
        self.acc_full = [MBar(f"acc_full{t}", 1) for t in range(TILES)]
        self.acc_empty = [MBar(f"acc_empty{t}", EPI_WARPS) for t in range(TILES)]
        self.stage_data = [None] * K_STAGES
        self.stage_copying = [False] * K_STAGES
        self.stage_readers = [0] * K_STAGES
        self.chunk_ver = [[None] * EPI_WARPS for _ in range(TILES * 4)]
        self.chunk_readers = [0] * (TILES * 4)
        self.acc_ver = [None] * TILES
        self.acc_writing = [None] * TILES
        self.acc_reads_left = [0] * TILES

    # ---- asynchronous agents (tile-indexed variants of the base class's)
    def mma(self, it, j, t, kc, stage, first):
        if self.stage_data[stage] != (it, j, t, kc):
            raise Hazard(f"MMA {(it, j, t, kc)} reads ring stage {stage} holding {self.stage_data[stage]}")
        phys = t * 4 + kc
        for w in range(EPI_WARPS):
            if self.chunk_ver[phys][w] != (it, j):
                raise Hazard(f"MMA {(it, j, t, kc)}: warp {w} wrote {self.chunk_ver[phys][w]} into chunk {phys}")
        if first:
            if self.acc_reads_left[t]:
                raise Hazard(f"step {(it, j, t)} overwrites its accumulator with {self.acc_reads_left[t]} reads outstanding")
            self.acc_writing[t] = (it, j)
            self.acc_ver[t] = None
        elif self.acc_writing[t] != (it, j):
            raise Hazard(f"accumulating {(it, j, t)} into an accumulator owned by {self.acc_writing[t]}")
        self.stage_readers[stage] += 1
        self.chunk_readers[phys] += 1
        start = max(self.now, self.mma_busy_until)
        self.mma_busy_until = start + self.lat(0.5, 2.0)

        def fin():
            self.stage_readers[stage] -= 1
            self.chunk_readers[phys] -= 1
        self.mma_queue.append((self.mma_busy_until, fin))
        self.at(self.mma_busy_until - self.now, self._retire)

    # ---- roles (mirroring mlp_rev2.cu)
    def producer(self):
        stage, rnd = 0, 0
        for it in range(self.iters):
            for jt in range(REV_LAYERS * TILES):
                j, t = jt >> 1, jt & 1
                for kc in range(4):
                    if rnd > 0:
                        yield ("wait", self.empty[stage], (rnd - 1) & 1)
                    self.full[stage].arrive(tx=1)
                    self.bulk_copy(stage, (it, j, t, kc))
                    yield ("delay", self.lat(0.05, 0.3))
                    stage += 1
                    if stage == K_STAGES:
                        stage, rnd = 0, rnd + 1

    def issuer(self):
        stage, rnd = 0, 0
        for it in range(self.iters):
            for jt in range(REV_LAYERS * TILES):
                j, t = jt >> 1, jt & 1
                started = it * 7 + j
                if started > 0:
                    yield ("wait", self.acc_empty[t], (started - 1) & 1)
                for kc in range(4):
                    yield ("wait", self.a_ready[t * 4 + kc], (it * self.a_per_tile + j) & 1)
                    yield ("wait", self.full[stage], rnd & 1)
                    self.mma(it, j, t, kc, stage, first=(kc == 0))
                    st = stage
                    self.commit(lambda st=st: self.empty[st].arrive())
                    yield ("delay", self.lat(0.05, 0.4))
                    stage += 1
                    if stage == K_STAGES:
                        stage, rnd = 0, rnd + 1

                def full_fn(t=t, it=it, j=j):
                    self.acc_ver[t] = (it, j)
                    self.acc_writing[t] = None
                    self.acc_reads_left[t] = EPI_WARPS
                    self.acc_full[t].arrive()
                self.commit(full_fn)
                yield ("delay", self.lat(0.02, 0.1))

    def epilogue(self, w):
        for it in range(self.iters):
            for j in range(-1, REV_LAYERS):
                lt = 6 - j
                for t in range(TILES):
                    if j >= 0:
                        yield ("wait", self.acc_full[t], (it * 7 + j) & 1)
                    for chunk in range(4):
                        if j >= 0 and self.acc_ver[t] != (it, j):
                            raise Hazard(f"warp {w} reads accumulator {t}: holds {self.acc_ver[t]}, want {(it, j)}")
                        yield ("delay", self.lat(0.1, 1.0))
                        if lt >= 1:
                            phys = t * 4 + chunk
                            if self.chunk_readers[phys]:
                                raise Hazard(f"warp {w} writes chunk {phys} under {self.chunk_readers[phys]} MMAs in flight")
                            self.chunk_ver[phys][w] = (it, j + 1)
                            self.a_ready[phys].arrive()
                    if j >= 0:
                        self.acc_reads_left[t] -= 1
                        self.acc_empty[t].arrive()


def test_rev2_protocol_no_deadlock_no_hazard():
    for seed in range(40):
        Rev2Sim(iters=3, seed=seed).run()


def test_rev2_protocol_under_heavy_tailed_latencies(monkeypatch):
    from tests.test_rg_protocol import _heavy_tailed
    monkeypatch.setattr(Sim, "lat", _heavy_tailed)
    for seed in range(300, 340):
        Rev2Sim(iters=4, seed=seed).run()


def test_rev2_model_detects_a_wrong_parity():
    with pytest.raises(AssertionError):
        for seed in range(10):
            Rev2Sim(iters=3, seed=seed, a_per_tile=8).run()
