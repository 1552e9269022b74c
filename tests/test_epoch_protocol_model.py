"""The epoch protocol of linked slabs, as a happens-before model (CPU; no GPU, no library call).

Neighbouring slabs of one lattice (a chain A above B above C ...) store into each other's halo rows from their own
streams and order those stores with one 64-bit epoch word per face.  A GPU soak with four processes time-slicing one
device found a hole in that protocol once in 400 random walks; the hole is a property of WHICH launches wait for and
publish epochs, not of any kernel, so it can be checked without a GPU: this file restates, call by call, the
wait / launch / signal sequence that lbm_b200/csrc/api.cu enqueues (sync_peers, signal_peers, push_all_halos,
launch_step, do_steps, materialise, run_summary, blbm_collide, blbm_stream, blbm_custom_speed, blbm_write_population +
blbm_exchange_halos), derives the happens-before relation the epochs establish between the two streams (a wait for
epoch e is released by the neighbour's first signal >= e; no assumption whatsoever about the slabs' relative speed),
and requires every pair of accesses to the same halo region from the two streams, one of them a store, to be ordered -
unless the store only re-sends what the region provably holds already (push_all_halos re-sends the boundary rows of
both buffers; the rows a pending gather of the neighbour reads have not changed since they were last pushed: every
store carries a version of the rows it copies, and a gather may race with a store of the version it would read anyway).

It is a model: the GPU tests (test_gpu_parity.py::test_a_slab_running_ahead_..., the group and multi-process tests)
check the real thing.  What the model adds is exhaustiveness - every call sequence up to a length, both handshake
flavours, read-backs on one slab alone - and the negative controls: with the epoch after the summary, or after the
public stream half-step, removed (the protocol before the fix) it reports exactly the races that were seen on the GPU;
a third one of the same family (write_population + exchange_halos against a neighbour's pending stream pass) was found
by the model itself and closed the same way.
"""
import itertools
import os

import pytest

FUSED, COLLIDE, STREAM = "fused", "collide-only", "stream-only"


class Slab:
    """host-side state of one slab handle and the operations its calls enqueue on its stream"""

    def __init__(self, name, up=None, dn=None, in_kernel=True, summary_epoch=True, stream_epoch=True,
                 exchange_epoch=True):
        self.name = name
        self.nbr = {side: n for side, n in (("up", up), ("dn", dn)) if n}  # the slabs above / below, if any
        self.in_kernel, self.summary_epoch, self.stream_epoch = in_kernel, summary_epoch, stream_epoch
        self.exchange_epoch = exchange_epoch
        self.regimeT, self.step, self.halo_dirty = False, 0, True  # blbm_link_*: halo_dirty = true
        self.epoch, self.waited = 0, 0
        self.ops = []  # ('wait', e) | ('signal', e) | ('kernel', what, reads {region}, writes {region: tag})
        self.ver = {("pop", 0): 0, ("pop", 1): 0, ("mom", None): 0}  # versions of our own boundary rows

    def touch(self, kind, buf=None):
        """a kernel of ours rewrites our own rows of that plane: what a later push copies is a new value"""
        self.ver[(kind, buf)] += 1

    # regions: (owner of the memory, which of its two halos, kind, buffer)
    def pushed(self, kind, buf=None):
        """regions in the neighbours' memory -> tag of the rows we copy there (our first row into the lower halo of
        the slab above, our last row into the upper halo of the slab below)"""
        return {(n, "dn" if side == "up" else "up", kind, buf): (self.name, side, kind, buf, self.ver[(kind, buf)])
                for side, n in self.nbr.items()}

    def own(self, kind, buf=None):
        """our halo rows that a gather of ours reads"""
        return {(self.name, side, kind, buf) for side in self.nbr}

    # ---- api.cu: sync_peers / signal_peers / push_all_halos --------------------------------------------------
    def sync_peers(self):
        if self.waited >= self.epoch:
            return
        self.ops.append(("wait", self.epoch))
        self.waited = self.epoch

    def signal_peers(self):
        self.epoch += 1
        self.ops.append(("signal", self.epoch))

    def push_all_halos(self):
        self.sync_peers()
        self.ops.append(("kernel", "push_all_halos", set(),
                         {**self.pushed("pop", 0), **self.pushed("pop", 1), **self.pushed("mom")}))
        self.halo_dirty = False
        self.signal_peers()

    # ---- api.cu: launch_step ---------------------------------------------------------------------------------
    def launch_step(self, mode, xbuf, ybuf, mom):
        pushes = mode != STREAM
        if self.halo_dirty and mode != COLLIDE:
            self.push_all_halos()
        reads = self.own("pop", xbuf) if mode in (FUSED, STREAM) else set()
        self.touch("pop", ybuf)  # every mode rewrites our rows of buffer y
        writes = dict(self.pushed("pop", ybuf)) if pushes else {}
        if pushes and mom:
            self.touch("mom")
            writes.update(self.pushed("mom"))
        kernel = ("kernel", f"{mode} x={xbuf} y={ybuf}{' +moments' if mom else ''}", reads, writes)
        if self.in_kernel and mode == FUSED:  # the fused vec4 kernel waits and publishes itself (face row blocks)
            self.ops += [("wait", self.epoch), kernel, ("signal", self.epoch + 1)]
            self.waited = self.epoch
            self.epoch += 1
            return
        self.sync_peers()
        self.ops.append(kernel)
        if mode == COLLIDE:
            self.halo_dirty = False
        if pushes:
            self.signal_peers()

    # ---- api.cu: do_steps / materialise / run_summary ---------------------------------------------------------
    def do_steps(self, n):
        for left in range(n, 0, -1):
            mom = left == 1
            if not self.regimeT:
                y = self.step % 2
                self.launch_step(COLLIDE, y, y, mom)
                self.regimeT = True
            else:
                self.launch_step(FUSED, (self.step + 1) % 2, self.step % 2, mom)
            self.step += 1

    def materialise(self):
        if not self.regimeT:
            return
        self.launch_step(STREAM, (self.step + 1) % 2, self.step % 2, False)
        self.regimeT = False
        self.halo_dirty = True

    def run_summary(self):
        self.sync_peers()
        self.ops.append(("kernel", "summary (curl)", self.own("mom"), {}))
        if self.summary_epoch:
            self.signal_peers()

    # ---- the C ABI calls -----------------------------------------------------------------------------------------
    def call(self, what):
        if what.startswith("iterate"):
            self.do_steps(int(what[7:]))
            self.run_summary()
        elif what.startswith("advance"):
            self.do_steps(int(what[7:]))
        elif what == "rerender":
            self.run_summary()
        elif what == "collide":
            self.materialise()
            y = self.step % 2
            self.launch_step(COLLIDE, y, y, True)
        elif what == "stream":
            self.materialise()
            self.launch_step(STREAM, self.step % 2, (self.step + 1) % 2, False)
            if self.stream_epoch:
                self.signal_peers()
        elif what == "reset":  # custom_speed / reset_to_equilibrium / single_cell: own rows AND own halo rows
            self.sync_peers()
            for key in self.ver:
                self.touch(*key)
            self.ops.append(("kernel", "fill", set(), {region: (self.name, "fill", key, self.ver[key])
                                                      for key in self.ver for region in self.own(*key)}))
            self.step, self.regimeT, self.halo_dirty = 0, False, False
            self.signal_peers()
        elif what == "readback":  # read_population & co: brings both buffers to the reference's state, own rows only
            self.materialise()
        elif what == "restore":  # write_population (checkpoint) followed by the collective exchange_halos
            self.materialise()
            self.touch("pop", 0)
            self.touch("pop", 1)
            self.halo_dirty = True
            self.materialise()  # blbm_exchange_halos
            if self.exchange_epoch:
                self.signal_peers()  # "everything I still had to gather from my halo rows is done"
            self.push_all_halos()
        else:
            raise ValueError(what)


def happens_before(slabs):
    """vector clocks of every operation of every stream; returns (clocks, deadlock) - a wait for epoch e is released
    once every neighbour's first signal with a value >= e has been issued"""
    names = list(slabs)
    clocks = {n: [] for n in names}
    done = {n: 0 for n in names}
    cur = {n: [0] * len(names) for n in names}
    progress = True
    while progress:
        progress = False
        for me in names:
            while done[me] < len(slabs[me].ops):
                op = slabs[me].ops[done[me]]
                if op[0] == "wait" and op[1] > 0:
                    sigs = {}
                    for other in slabs[me].nbr.values():
                        sigs[other] = next((k for k, o in enumerate(slabs[other].ops)
                                            if o[0] == "signal" and o[1] >= op[1]), None)
                    if any(k is None or k >= done[o] for o, k in sigs.items()):
                        break  # not published yet
                    for o, k in sigs.items():
                        cur[me] = [max(x, y) for x, y in zip(cur[me], clocks[o][k])]
                cur[me] = list(cur[me])
                cur[me][names.index(me)] += 1
                clocks[me].append(tuple(cur[me]))
                done[me] += 1
                progress = True
    deadlock = any(done[n] < len(slabs[n].ops) for n in names)
    return clocks, deadlock


def races(slabs):
    """accesses of different streams to the same halo region, at least one of them a store, with no happens-before
    between them - except a store that re-sends the version the region already holds for the racing gather"""
    clocks, deadlock = happens_before(slabs)
    if deadlock:
        return ["deadlock"]

    def hb(p, q):  # p happens before q
        return p != q and all(x <= y for x, y in zip(p, q))

    reads, writes = [], []
    for s, slab in slabs.items():
        for k, op in enumerate(slab.ops):
            if op[0] == "kernel":
                reads += [(region, s, clocks[s][k], op[1]) for region in op[2]]
                writes += [(region, tag, s, clocks[s][k], op[1]) for region, tag in op[3].items()]
    found = []
    for (r1, t1, s1, c1, w1), (r2, t2, s2, c2, w2) in itertools.combinations(writes, 2):
        if r1 == r2 and s1 != s2 and t1 != t2 and not hb(c1, c2) and not hb(c2, c1):
            found.append(f"{r1}: {s1} w [{w1}] || {s2} w [{w2}]")
    for region, s, c, what in reads:
        mine = [w for w in writes if w[0] == region]
        racing = [w for w in mine if w[2] != s and not hb(w[3], c) and not hb(c, w[3])]
        if not racing:
            continue
        before = [w for w in mine if hb(w[3], c)]
        last = [w for w in before if not any(hb(w[3], v[3]) for v in before)]  # what the region holds for this gather
        held = {w[1] for w in last}
        for w in racing:
            if len(held) != 1 or w[1] not in held:
                found.append(f"{region}: {s} r [{what}] || {w[2]} w [{w[4]}]")
    return found


def chain(n, **kw):
    names = "ABCDE"[:n]
    return {c: Slab(c, up=names[i - 1] if i else None, dn=names[i + 1] if i + 1 < n else None, **kw)
            for i, c in enumerate(names)}


def run(calls_a, calls_b=None, nslabs=2, **kw):
    """slab A is given calls_a, every other slab calls_b (default: the same calls)"""
    slabs = chain(nslabs, **kw)
    for name, slab in slabs.items():
        for c in (calls_a if name == "A" or calls_b is None else calls_b):
            slab.call(c)
    return races(slabs)


CALLS = ["iterate1", "iterate2", "iterate3", "advance1", "advance2", "collide", "stream", "rerender", "reset",
         "readback", "restore"]


DEPTH = int(os.environ.get("BLBM_PROTOCOL_MODEL_DEPTH", "4"))  # 5: 177,155 sequences, a few minutes (soak)


@pytest.mark.parametrize("in_kernel", [True, False])
def test_every_call_sequence_up_to_four_calls_is_race_free(in_kernel):
    """both slabs are given the same calls (include/blbm.h: the rule for linked slabs); any relative speed"""
    n = 0
    for length in range(1, DEPTH + 1):
        for seq in itertools.product(CALLS, repeat=length):
            found = run(seq, in_kernel=in_kernel)
            assert not found, f"{seq}: {found[:3]}"
            n += 1
    assert n == sum(len(CALLS) ** k for k in range(1, DEPTH + 1))


@pytest.mark.parametrize("in_kernel", [True, False])
@pytest.mark.parametrize("nslabs", [2, 3, 4])
def test_long_random_call_sequences_with_stray_read_backs_are_race_free(in_kernel, nslabs):
    """chains of 2-4 slabs (a middle slab waits for and publishes to two neighbours), 5-12 calls, read-backs strewn
    over each slab independently"""
    import random
    rng = random.Random(20230 + nslabs)
    base = [c for c in CALLS if c != "readback"]
    for _ in range(600):
        seq = [rng.choice(base) for _ in range(rng.randint(5, 12))]
        slabs = chain(nslabs, in_kernel=in_kernel)
        calls = {}
        for name, slab in slabs.items():
            calls[name] = list(seq)
            for _ in range(rng.randint(0, 3)):  # read-backs wherever a rank likes, independently of its neighbours
                calls[name].insert(rng.randint(0, len(calls[name])), "readback")
            for c in calls[name]:
                slab.call(c)
        found = races(slabs)
        assert not found, f"{calls}: {found[:3]}"


@pytest.mark.parametrize("in_kernel", [True, False])
def test_every_call_sequence_up_to_three_calls_on_three_slabs(in_kernel):
    for length in (1, 2, 3):
        for seq in itertools.product(CALLS, repeat=length):
            found = run(seq, nslabs=3, in_kernel=in_kernel)
            assert not found, f"{seq}: {found[:3]}"


@pytest.mark.parametrize("in_kernel", [True, False])
def test_read_backs_on_one_slab_alone_stay_legal(in_kernel):
    """materialise() publishes no epoch: a read-back (which runs the pending stream pass) on one slab only neither
    desynchronises the epochs nor opens a race, wherever it falls in the call sequence"""
    base = [c for c in CALLS if c != "readback"]
    for seq in itertools.product(base, repeat=3):
        for at in range(4):
            skewed = list(seq[:at]) + ["readback"] + list(seq[at:])
            for calls_a, calls_b in ((skewed, list(seq)), (list(seq), skewed)):
                found = run(calls_a, calls_b, in_kernel=in_kernel)
                assert not found, f"A {calls_a} / B {calls_b}: {found[:3]}"


@pytest.mark.parametrize("in_kernel", [True, False])
def test_the_model_reports_the_races_of_the_protocol_before_the_fixes(in_kernel):
    """negative controls: the protocol before the fixes"""
    # 1. no epoch after the summary: the neighbour's next moment-storing launch against the pending curl
    for seq in (("iterate1", "collide"), ("iterate3", "iterate1"), ("rerender", "collide")):
        found = run(seq, in_kernel=in_kernel, summary_epoch=False)
        assert any("'mom'" in f and "summary" in f for f in found), (seq, found)
    assert not run(("iterate3", "iterate2"), in_kernel=in_kernel, summary_epoch=False)  # two steps later it was safe
    # 2. no epoch after the public stream half-step: the neighbour's next collide against the pending gather
    found = run(("iterate2", "stream", "collide"), in_kernel=in_kernel, stream_epoch=False)
    assert any("'pop'" in f and "stream-only" in f and "collide-only" in f for f in found), found
    # 3. (found by this model, not on the GPU) exchange_halos without an epoch of its own: rows rewritten by
    #    write_population pushed into halo rows that the neighbour's pending stream pass still gathers from
    found = run(("iterate1", "restore"), in_kernel=in_kernel, exchange_epoch=False)
    assert any("'pop'" in f and "stream-only" in f and "push_all_halos" in f for f in found), found
    # a waiter without a matching signal is reported, not looped on
    slabs = chain(2)
    slabs["A"].call("iterate2")
    assert races(slabs) == ["deadlock"]
